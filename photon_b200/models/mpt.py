"""MPT decoder-only LM (``mpt_causal_lm``) in plain PyTorch.

This is the *oracle / CPU / baseline* implementation: parameter names, init
and math follow what the reference trains through llm-foundry
(ref: photon/conf/llm_config/mpt-125m.yaml:18-28; names relied on at
photon/utils.py:572,591,602-637). The B200 hot path lives in
:mod:`photon_b200.models.engine`, which runs the same parameters through
hand-written sm_100a kernels and is tested against this module.

Block: ``x = x + attn(norm_1(x)); x = x + ffn(norm_2(x))``; ``ffn =
down(gelu_exact(up(x)))``; LM head tied to ``wte``; loss = CE over
``targets = roll(labels, -1)`` with the last position ignored (-100).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class MPTConfig:
    d_model: int = 768
    n_heads: int = 12
    n_layers: int = 12
    expansion_ratio: int = 4
    max_seq_len: int = 2048
    vocab_size: int = 50368
    attn_impl: str = "flash"  # flash | torch | b200
    alibi: bool = False
    alibi_bias_max: int = 8
    rope: bool = False
    rope_theta: float = 10000.0
    qk_ln: bool = False
    clip_qkv: float | None = None
    no_bias: bool = False
    learned_pos_emb: bool = True
    norm_eps: float = 1e-5
    init_std: float | None = None  # None -> kaiming_normal_ (llm-foundry default)
    extra: dict[str, Any] = field(default_factory=dict)

    @property
    def d_head(self) -> int:
        return self.d_model // self.n_heads

    @classmethod
    def from_model_cfg(cls, m: dict[str, Any]) -> "MPTConfig":
        """Build from ``llm_config.model`` (unknown llm-foundry keys are kept in ``extra``)."""
        m = dict(m)
        attn = dict(m.pop("attn_config", {}) or {})
        rope_cfg = attn.pop("rope_impl", None)
        known = dict(
            d_model=m.pop("d_model"), n_heads=m.pop("n_heads"), n_layers=m.pop("n_layers"),
            expansion_ratio=m.pop("expansion_ratio", 4), max_seq_len=m.pop("max_seq_len", 2048),
            vocab_size=m.pop("vocab_size", 50368), attn_impl=attn.pop("attn_impl", "flash"),
            alibi=bool(attn.pop("alibi", False)), alibi_bias_max=attn.pop("alibi_bias_max", 8),
            rope=bool(attn.pop("rope", False)), rope_theta=float(attn.pop("rope_theta", 10000.0)),
            qk_ln=bool(attn.pop("qk_ln", False)), clip_qkv=attn.pop("clip_qkv", None),
            no_bias=bool(m.pop("no_bias", False)), learned_pos_emb=bool(m.pop("learned_pos_emb", True)),
        )
        for k in ("name", "init_device"):
            m.pop(k, None)
        cfg = cls(**known, extra={**m, **({"attn_config": attn} if attn else {}),
                                  **({"rope_impl": rope_cfg} if rope_cfg else {})})
        if cfg.d_model % cfg.n_heads:
            raise ValueError("d_model must be divisible by n_heads")
        if cfg.alibi or cfg.rope:
            cfg.learned_pos_emb = False if (cfg.alibi or cfg.rope) and "learned_pos_emb" not in m else cfg.learned_pos_emb
        return cfg

    def num_params(self) -> int:
        d, L, V, S, e = self.d_model, self.n_layers, self.vocab_size, self.max_seq_len, self.expansion_ratio
        b = 0 if self.no_bias else 1
        per_block = (2 * d) * 2 + (3 * d * d + b * 3 * d) + (d * d + b * d) + (e * d * d + b * e * d) + (e * d * d + b * d)
        if self.no_bias:
            per_block -= 2 * d  # LN biases dropped too
        return V * d + (S * d if self.learned_pos_emb else 0) + L * per_block + (2 * d if not self.no_bias else d)

    def flops_per_token(self, seq_len: int | None = None) -> float:
        """fwd+bwd model FLOPs per token as SpeedMonitor counts them: 6·N + 12·L·d·S."""
        s = seq_len or self.max_seq_len
        return 6.0 * self.num_params() + 12.0 * self.n_layers * self.d_model * s


def alibi_slopes(n_heads: int, bias_max: int = 8) -> torch.Tensor:
    """Per-head ALiBi slopes (power-of-two ladder, interleaved for non-pow2 head counts)."""
    p2 = 2 ** math.ceil(math.log2(n_heads))
    m = torch.arange(1, p2 + 1, dtype=torch.float32) * (bias_max / p2)
    slopes = 1.0 / torch.pow(2.0, m)
    if p2 != n_heads:
        slopes = torch.cat([slopes[1::2], slopes[::2]])[:n_heads]
    return slopes


def rope_tables(seq_len: int, d_head: int, theta: float, device: Any = None) -> tuple[torch.Tensor, torch.Tensor]:
    inv = 1.0 / (theta ** (torch.arange(0, d_head, 2, dtype=torch.float32, device=device) / d_head))
    ang = torch.outer(torch.arange(seq_len, dtype=torch.float32, device=device), inv)
    return ang.cos(), ang.sin()


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x: [B,H,S,dh]; rotate-half (GPT-NeoX) convention."""
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    c, s = cos[None, None].to(x.dtype), sin[None, None].to(x.dtype)
    return torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1)


class MPTAttention(nn.Module):
    def __init__(self, cfg: MPTConfig) -> None:
        super().__init__()
        self.cfg = cfg
        d = cfg.d_model
        self.Wqkv = nn.Linear(d, 3 * d, bias=not cfg.no_bias)
        self.out_proj = nn.Linear(d, d, bias=not cfg.no_bias)
        if cfg.qk_ln:
            self.q_ln = nn.LayerNorm(d, eps=cfg.norm_eps)
            self.k_ln = nn.LayerNorm(d, eps=cfg.norm_eps)

    def forward(self, x: torch.Tensor, rope: tuple[torch.Tensor, torch.Tensor] | None,
                alibi: torch.Tensor | None) -> torch.Tensor:
        cfg = self.cfg
        B, S, d = x.shape
        qkv = self.Wqkv(x)
        if cfg.clip_qkv:
            qkv = qkv.clamp(-cfg.clip_qkv, cfg.clip_qkv)
        q, k, v = qkv.chunk(3, dim=-1)
        if cfg.qk_ln:
            q, k = self.q_ln(q).to(q.dtype), self.k_ln(k).to(k.dtype)
        H, dh = cfg.n_heads, cfg.d_head
        q, k, v = (t.view(B, S, H, dh).transpose(1, 2) for t in (q, k, v))
        if rope is not None:
            q, k = apply_rope(q, rope[0][:S], rope[1][:S]), apply_rope(k, rope[0][:S], rope[1][:S])
        scale = 1.0 / math.sqrt(dh)
        if cfg.attn_impl == "torch" or (alibi is not None and cfg.attn_impl != "flash"):
            att = (q @ k.transpose(-1, -2)) * scale
            if alibi is not None:
                att = att + alibi[:, :, :S, :S].to(att.dtype)
            mask = torch.ones(S, S, dtype=torch.bool, device=x.device).tril()
            att = att.masked_fill(~mask, float("-inf"))
            o = torch.softmax(att.float(), dim=-1).to(v.dtype) @ v
        else:
            bias = None
            if alibi is not None:
                mask = torch.ones(S, S, dtype=torch.bool, device=x.device).tril()
                bias = alibi[:, :, :S, :S].to(q.dtype).masked_fill(~mask, float("-inf"))
            o = F.scaled_dot_product_attention(q, k, v, attn_mask=bias, is_causal=bias is None, scale=scale)
        return self.out_proj(o.transpose(1, 2).reshape(B, S, d))


class MPTMLP(nn.Module):
    def __init__(self, cfg: MPTConfig) -> None:
        super().__init__()
        d = cfg.d_model
        self.up_proj = nn.Linear(d, cfg.expansion_ratio * d, bias=not cfg.no_bias)
        self.down_proj = nn.Linear(cfg.expansion_ratio * d, d, bias=not cfg.no_bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.down_proj(F.gelu(self.up_proj(x)))


class MPTBlock(nn.Module):
    def __init__(self, cfg: MPTConfig) -> None:
        super().__init__()
        self.norm_1 = nn.LayerNorm(cfg.d_model, eps=cfg.norm_eps, bias=not cfg.no_bias)
        self.attn = MPTAttention(cfg)
        self.norm_2 = nn.LayerNorm(cfg.d_model, eps=cfg.norm_eps, bias=not cfg.no_bias)
        self.ffn = MPTMLP(cfg)

    def forward(self, x: torch.Tensor, rope: Any, alibi: Any) -> torch.Tensor:
        x = x + self.attn(self.norm_1(x).to(x.dtype), rope, alibi)
        return x + self.ffn(self.norm_2(x).to(x.dtype))


class MPTTransformer(nn.Module):
    def __init__(self, cfg: MPTConfig) -> None:
        super().__init__()
        self.wte = nn.Embedding(cfg.vocab_size, cfg.d_model)
        if cfg.learned_pos_emb:
            self.wpe = nn.Embedding(cfg.max_seq_len, cfg.d_model)
        self.blocks = nn.ModuleList([MPTBlock(cfg) for _ in range(cfg.n_layers)])
        self.norm_f = nn.LayerNorm(cfg.d_model, eps=cfg.norm_eps, bias=not cfg.no_bias)


class MPTForCausalLM(nn.Module):
    """``transformer.*`` parameter tree + tied LM head + shifted-target CE loss."""

    def __init__(self, cfg: MPTConfig, device: Any = None, init: bool = True, seed: int | None = None) -> None:
        super().__init__()
        self.cfg = cfg
        self.activation_checkpointing = False   # fsdp_config.activation_checkpointing (ref: conf/llm_config/mpt-1b.yaml:88)
        with torch.device(device or "cpu"):
            self.transformer = MPTTransformer(cfg)
        if init:
            self.reset_parameters(seed)
        self._rope_cache: tuple[torch.Tensor, torch.Tensor] | None = None
        self._alibi_cache: torch.Tensor | None = None

    # -- init (llm-foundry ``kaiming_normal_`` param_init_fn) ----
    @torch.no_grad()
    def reset_parameters(self, seed: int | None = None) -> None:
        cfg = self.cfg
        gen = None
        if seed is not None:
            gen = torch.Generator(device="cpu").manual_seed(int(seed))
        div = math.sqrt(2 * cfg.n_layers)

        def _fill(w: torch.Tensor, fan_in: int, residual: bool = False) -> None:
            std = cfg.init_std if cfg.init_std is not None else math.sqrt(2.0 / fan_in)
            tmp = torch.empty(w.shape, dtype=torch.float32).normal_(0.0, std, generator=gen)
            if residual:
                tmp /= div
            w.copy_(tmp.to(w.device, w.dtype))

        for name, mod in self.named_modules():
            if isinstance(mod, nn.Linear):
                if name.endswith("Wqkv"):  # fused: initialise q/k/v slices independently
                    d = cfg.d_model
                    for s in range(3):
                        _fill(mod.weight[s * d:(s + 1) * d], mod.in_features)
                else:
                    _fill(mod.weight, mod.in_features, residual=name.endswith(("out_proj", "down_proj")))
                if mod.bias is not None:
                    mod.bias.zero_()
            elif isinstance(mod, nn.Embedding):
                _fill(mod.weight, mod.embedding_dim)
            elif isinstance(mod, nn.LayerNorm):
                mod.weight.fill_(1.0)
                if mod.bias is not None:
                    mod.bias.zero_()

    # -- forward ----------------------------------------------------------------
    def _aux(self, S: int, device: torch.device) -> tuple[Any, Any]:
        cfg = self.cfg
        rope = alibi = None
        if cfg.rope:
            if self._rope_cache is None or self._rope_cache[0].device != device:
                self._rope_cache = rope_tables(cfg.max_seq_len, cfg.d_head, cfg.rope_theta, device)
            rope = self._rope_cache
        if cfg.alibi:
            if self._alibi_cache is None or self._alibi_cache.device != device:
                sl = alibi_slopes(cfg.n_heads, cfg.alibi_bias_max).to(device)
                pos = torch.arange(cfg.max_seq_len, device=device)
                rel = (pos[None, :] - pos[:, None]).clamp(max=0).float()  # -(i-j) for j<=i
                self._alibi_cache = (sl[:, None, None] * rel[None])[None]
            alibi = self._alibi_cache
        return rope, alibi

    def hidden_states(self, input_ids: torch.Tensor) -> torch.Tensor:
        t = self.transformer
        B, S = input_ids.shape
        if S > self.cfg.max_seq_len:
            raise ValueError(f"sequence length {S} > max_seq_len {self.cfg.max_seq_len}")
        x = t.wte(input_ids)
        if self.cfg.learned_pos_emb:
            x = x + t.wpe(torch.arange(S, device=input_ids.device))[None]
        rope, alibi = self._aux(S, input_ids.device)
        for blk in t.blocks:
            if self.activation_checkpointing and self.training and torch.is_grad_enabled():
                from torch.utils.checkpoint import checkpoint

                x = checkpoint(blk, x, rope, alibi, use_reentrant=False)   # keep only block inputs, recompute in backward
            else:
                x = blk(x, rope, alibi)
        return t.norm_f(x).to(x.dtype)

    def forward(self, input_ids: torch.Tensor) -> torch.Tensor:
        h = self.hidden_states(input_ids)
        return F.linear(h, self.transformer.wte.weight.to(h.dtype))

    def loss(self, input_ids: torch.Tensor, labels: torch.Tensor | None = None,
             reduction: str = "mean") -> tuple[torch.Tensor, torch.Tensor]:
        """Returns (loss, n_valid_tokens). Targets are shifted *inside* (ref semantics)."""
        logits = self.forward(input_ids)
        targets = shift_labels(input_ids if labels is None else labels)
        flat = F.cross_entropy(logits.float().view(-1, logits.shape[-1]), targets.view(-1),
                               ignore_index=-100, reduction="sum")
        n = (targets != -100).sum()
        return (flat / n.clamp(min=1) if reduction == "mean" else flat), n


def shift_labels(labels: torch.Tensor) -> torch.Tensor:
    """``targets = roll(labels, -1); targets[:, -1] = -100``."""
    t = torch.roll(labels, shifts=-1, dims=1).clone()
    t[:, -1] = -100
    return t


def build_model(model_cfg: dict[str, Any] | MPTConfig, device: Any = None, seed: int | None = None,
                init: bool = True) -> MPTForCausalLM:
    cfg = model_cfg if isinstance(model_cfg, MPTConfig) else MPTConfig.from_model_cfg(model_cfg)
    return MPTForCausalLM(cfg, device=device, init=init, seed=seed)


def trainable_named_parameters(model: nn.Module) -> list[tuple[str, nn.Parameter]]:
    """Trainable parameters in **lexicographic name order** — the order baked into
    every exchanged payload and ``.npz`` (``blocks.10`` < ``blocks.2``; ref:
    photon/utils.py:316-317, photon/clients/utils.py:854-857)."""
    return sorted(((n, p) for n, p in model.named_parameters() if p.requires_grad), key=lambda kv: kv[0])
