"""Config composer: a small, dependency-free stand-in for the Hydra/OmegaConf
features the reference relies on (ref: photon/hydra_resolver.py:11-39,
photon/conf/base.yaml:76-83, scripts/fed_125m_example.sh:46-103).

Supported grammar
-----------------
* ``defaults`` lists in any YAML file: ``_self_``, ``group: option``,
  ``group@pkg.path: option`` and ``optional group: option``; nested defaults
  inside group files are honoured (used by ``llm_config/mpt-*.yaml`` here).
* overrides: ``a.b=1`` (key must exist), ``+a.b=1`` (must NOT exist),
  ``++a.b=1`` (force), ``~a.b`` (delete), ``group=option`` and
  ``group@pkg=option`` (swap a defaults entry), ``hydra/...`` (ignored).
* ``${a.b.c}`` interpolation, whole-node (typed) or inside strings, plus
  ``${oc.env:VAR,default}``.

The result is a :class:`ConfigNode` (a dict with attribute access) so call
sites read ``cfg.fl.n_rounds`` exactly like the reference's DictConfig.
"""
from __future__ import annotations

import copy
import os
import re
from pathlib import Path
from typing import Any, Iterable

import yaml

_MISSING = object()


class ConfigError(ValueError):
    """Raised for any composition / override / interpolation problem."""


class ConfigNode(dict):
    """Dict with attribute access, recursive wrapping and dotted-path helpers."""

    def __init__(self, data: dict | None = None) -> None:
        super().__init__()
        for k, v in (data or {}).items():
            self[k] = v

    @staticmethod
    def _wrap(v: Any) -> Any:
        if isinstance(v, ConfigNode):
            return v
        if isinstance(v, dict):
            return ConfigNode(v)
        if isinstance(v, (list, tuple)):
            return [ConfigNode._wrap(x) for x in v]
        return v

    def __setitem__(self, k: str, v: Any) -> None:
        super().__setitem__(k, self._wrap(v))

    def __getattr__(self, k: str) -> Any:
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k: str, v: Any) -> None:
        self[k] = v

    def __delattr__(self, k: str) -> None:
        try:
            del self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __deepcopy__(self, memo: dict) -> "ConfigNode":
        return ConfigNode({k: copy.deepcopy(v, memo) for k, v in self.items()})

    # dotted-path API --------------------------------------------------------
    def select(self, path: str, default: Any = _MISSING) -> Any:
        cur: Any = self
        for part in _split_path(path):
            if isinstance(cur, dict) and part in cur:
                cur = cur[part]
            elif isinstance(cur, list) and part.lstrip("-").isdigit() and -len(cur) <= int(part) < len(cur):
                cur = cur[int(part)]
            else:
                if default is _MISSING:
                    raise ConfigError(f"config key '{path}' not found (at '{part}')")
                return default
        return cur

    def has(self, path: str) -> bool:
        return self.select(path, default=_MISSING_SENTINEL) is not _MISSING_SENTINEL

    def update_path(self, path: str, value: Any, *, create: bool = True) -> None:
        parts = _split_path(path)
        cur: Any = self
        for part in parts[:-1]:
            if isinstance(cur, list):
                cur = cur[int(part)]
                continue
            if part not in cur or cur[part] is None:
                if not create:
                    raise ConfigError(f"config key '{path}' not found (at '{part}')")
                cur[part] = ConfigNode()
            cur = cur[part]
            if not isinstance(cur, (dict, list)):
                raise ConfigError(f"cannot descend into scalar at '{part}' for '{path}'")
        last = parts[-1]
        if isinstance(cur, list):
            cur[int(last)] = ConfigNode._wrap(value)
        else:
            cur[last] = value

    def delete_path(self, path: str) -> None:
        parts = _split_path(path)
        parent = self.select(".".join(parts[:-1])) if len(parts) > 1 else self
        if not isinstance(parent, dict) or parts[-1] not in parent:
            raise ConfigError(f"cannot delete missing key '{path}'")
        del parent[parts[-1]]

    def to_container(self) -> dict:
        return to_container(self)


_MISSING_SENTINEL = object()


def _split_path(path: str) -> list[str]:
    return [p for p in path.split(".") if p != ""]


def to_container(v: Any) -> Any:
    """Recursively convert ConfigNode/list trees into plain dict/list."""
    if isinstance(v, dict):
        return {k: to_container(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [to_container(x) for x in v]
    return v


def deep_merge(dst: dict, src: dict) -> dict:
    """Merge ``src`` into ``dst`` in place (dicts merge, everything else replaces)."""
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            deep_merge(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)
    return dst


# ----------------------------------------------------------------------------
# YAML IO
# ----------------------------------------------------------------------------
class _Loader(yaml.SafeLoader):
    """SafeLoader that reads 1e-6 / 6.0e-4 style floats like OmegaConf does."""


_Loader.add_implicit_resolver(
    "tag:yaml.org,2002:float",
    re.compile(
        r"""^(?:[-+]?(?:[0-9][0-9_]*)\.[0-9_]*(?:[eE][-+]?[0-9]+)?
        |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
        |\.[0-9_]+(?:[eE][-+][0-9]+)?
        |[-+]?\.(?:inf|Inf|INF)
        |\.(?:nan|NaN|NAN))$""",
        re.X,
    ),
    list("-+0123456789."),
)


def load_yaml(path: str | os.PathLike) -> Any:
    with open(path, "r", encoding="utf-8") as f:
        return yaml.load(f, Loader=_Loader)  # noqa: S506 - SafeLoader subclass


def parse_value(text: str) -> Any:
    """Parse an override right-hand side with YAML semantics (null, [..], {..})."""
    if text == "":
        return ""
    try:
        return yaml.load(text, Loader=_Loader)  # noqa: S506
    except yaml.YAMLError:
        return text


def save_yaml(cfg: Any, path: str | os.PathLike) -> None:
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    with open(path, "w", encoding="utf-8") as f:
        yaml.safe_dump(to_container(cfg), f, sort_keys=False, default_flow_style=False)


def load_config(path: str | os.PathLike) -> ConfigNode:
    """Load an already-resolved config (what every non-resolver process does;
    ref: photon/server_app.py:116, photon/centralised_train.py:62)."""
    data = load_yaml(path)
    if not isinstance(data, dict):
        raise ConfigError(f"{path}: top level must be a mapping")
    return ConfigNode(data)


# ----------------------------------------------------------------------------
# defaults-list composition
# ----------------------------------------------------------------------------
class _DefaultEntry:
    __slots__ = ("group", "package", "option", "optional", "is_self")

    def __init__(self, group: str = "", package: str | None = None, option: Any = None,
                 optional: bool = False, is_self: bool = False) -> None:
        self.group, self.package, self.option = group, package, option
        self.optional, self.is_self = optional, is_self

    @property
    def key(self) -> str:
        return self.group if self.package is None else f"{self.group}@{self.package}"


def _parse_defaults(raw: Iterable[Any], where: str) -> list[_DefaultEntry]:
    out: list[_DefaultEntry] = []
    for item in raw or []:
        if item == "_self_":
            out.append(_DefaultEntry(is_self=True))
            continue
        if isinstance(item, str):  # bare file include "group/option"
            grp, _, opt = item.rpartition("/")
            out.append(_DefaultEntry(group=grp, option=opt))
            continue
        if not (isinstance(item, dict) and len(item) == 1):
            raise ConfigError(f"{where}: bad defaults entry {item!r}")
        (k, opt), = item.items()
        optional = False
        if k.startswith("optional "):
            optional, k = True, k[len("optional "):].strip()
        if k.startswith("override "):
            k = k[len("override "):].strip()
        grp, _, pkg = k.partition("@")
        out.append(_DefaultEntry(group=grp, package=pkg or None, option=opt, optional=optional))
    if not any(e.is_self for e in out):
        out.insert(0, _DefaultEntry(is_self=True))
    return out


def _package_wrap(package: str, body: Any) -> dict:
    node: Any = body
    for part in reversed(_split_path(package)):
        node = {part: node}
    return node


def _compose_file(config_dir: Path, rel: str, choices: dict[str, Any], package: str,
                  used: dict[str, Any]) -> dict:
    """Compose one YAML file (and its defaults) into a dict rooted at ``package``."""
    path = config_dir / (rel + ".yaml")
    if not path.exists():
        raise ConfigError(f"config file not found: {path}")
    raw = load_yaml(path)
    if raw is None:
        raw = {}
    result: dict = {}
    if isinstance(raw, dict) and "defaults" in raw:
        body = {k: v for k, v in raw.items() if k != "defaults"}
        entries = _parse_defaults(raw["defaults"], str(path))
    else:
        body, entries = raw, [_DefaultEntry(is_self=True)]
    if isinstance(body, dict):  # "_anchor" helper keys never reach the tree
        body = {k: v for k, v in body.items() if not str(k).startswith("_")}
    parent_group = str(Path(rel).parent) if "/" in rel else ""
    for e in entries:
        if e.is_self:
            piece = _package_wrap(package, body) if package else body
            if isinstance(piece, dict):
                deep_merge(result, piece)
            else:  # a list-valued group file (e.g. dataset/streams/*.yaml)
                return piece  # type: ignore[return-value]
            continue
        if e.group == "" and isinstance(e.option, str):
            # bare include ("- _mpt_common"): sibling file merged at the SAME package
            sib = f"{parent_group}/{e.option}" if parent_group else e.option
            sub = _compose_file(config_dir, sib, choices, package, used)
            if isinstance(sub, dict):
                deep_merge(result, sub)
            continue
        # group path is relative to the including file's directory unless absolute
        grp = e.group[1:] if e.group.startswith("/") else (
            f"{parent_group}/{e.group}" if parent_group and not (config_dir / e.group).exists() else e.group)
        option = choices.get(e.key, choices.get(grp if e.package is None else f"{grp}@{e.package}", e.option))
        used[e.key] = option
        if option is None:
            continue
        sub_pkg_default = grp.replace("/", ".")
        if e.package is not None:
            sub_pkg = e.package
        elif package and parent_group and grp.startswith(parent_group + "/"):
            sub_pkg = package + "." + grp[len(parent_group) + 1:].replace("/", ".")
        else:
            sub_pkg = sub_pkg_default
        sub_rel = f"{grp}/{option}"
        if not (config_dir / (sub_rel + ".yaml")).exists():
            if e.optional:
                continue
            raise ConfigError(f"{path}: defaults entry '{e.key}: {option}' -> missing {sub_rel}.yaml")
        sub = _compose_file(config_dir, sub_rel, choices, sub_pkg, used)
        if isinstance(sub, dict):
            deep_merge(result, sub)
        else:
            deep_merge(result, _package_wrap(sub_pkg, sub))
    return result


# ----------------------------------------------------------------------------
# overrides
# ----------------------------------------------------------------------------
_OVR = re.compile(r"^(?P<prefix>\+\+|\+|~)?(?P<key>[^=]+?)(?:=(?P<val>.*))?$", re.S)


def _is_group(config_dir: Path, key: str) -> bool:
    grp = key.partition("@")[0]
    return (config_dir / grp).is_dir()


def split_overrides(config_dir: Path, overrides: Iterable[str]) -> tuple[dict[str, Any], list[tuple[str, str, Any]]]:
    """Separate group selections from value overrides (keeps CLI order)."""
    choices: dict[str, Any] = {}
    values: list[tuple[str, str, Any]] = []
    for ov in overrides:
        ov = ov.strip()
        if not ov:
            continue
        m = _OVR.match(ov)
        if not m:
            raise ConfigError(f"cannot parse override {ov!r}")
        prefix, key, val = m.group("prefix") or "", m.group("key").strip(), m.group("val")
        if key.startswith("hydra/") or key.startswith("hydra."):
            continue  # hydra/job_logging=none etc. (ref: scripts/photon_llm_125M.sh:129)
        if prefix == "~":
            values.append(("~", key, None))
            continue
        if val is None:
            raise ConfigError(f"override {ov!r} needs '=value'")
        if prefix == "" and _is_group(config_dir, key) and "/" not in val and "{" not in val:
            choices[key] = parse_value(val)
            continue
        values.append((prefix, key, parse_value(val)))
    return choices, values


def apply_overrides(cfg: ConfigNode, values: list[tuple[str, str, Any]]) -> None:
    for prefix, key, val in values:
        exists = cfg.select(key, default=_MISSING_SENTINEL) is not _MISSING_SENTINEL
        if prefix == "~":
            if exists:
                cfg.delete_path(key)
            else:
                raise ConfigError(f"'~{key}': key does not exist")
        elif prefix == "+":
            if exists:
                raise ConfigError(f"'+{key}': key already exists (use '++' to force)")
            cfg.update_path(key, val)
        elif prefix == "++":
            cfg.update_path(key, val)
        else:
            if not exists:
                raise ConfigError(f"'{key}': key not in config (use '+{key}=' to add it)")
            cfg.update_path(key, val)


# ----------------------------------------------------------------------------
# interpolation
# ----------------------------------------------------------------------------
_INTERP = re.compile(r"\$\{([^${}]+)\}")


def _resolve_ref(root: ConfigNode, ref: str, stack: tuple[str, ...]) -> Any:
    ref = ref.strip()
    if ref.startswith("oc.env:"):
        name, _, default = ref[len("oc.env:"):].partition(",")
        if name in os.environ:
            return parse_value(os.environ[name])
        if _ == "":
            raise ConfigError(f"environment variable '{name}' not set")
        return parse_value(default.strip())
    if ref in stack:
        raise ConfigError("interpolation cycle: " + " -> ".join(stack + (ref,)))
    target = root.select(ref, default=_MISSING_SENTINEL)
    if target is _MISSING_SENTINEL:
        raise ConfigError(f"interpolation '${{{ref}}}' not found")
    return _resolve_node(root, copy.deepcopy(target), stack + (ref,))


def _resolve_node(root: ConfigNode, node: Any, stack: tuple[str, ...] = ()) -> Any:
    if isinstance(node, dict):
        for k in list(node.keys()):
            node[k] = _resolve_node(root, node[k], stack)
        return node
    if isinstance(node, list):
        return [_resolve_node(root, x, stack) for x in node]
    if isinstance(node, str) and "${" in node:
        whole = _INTERP.fullmatch(node.strip())
        if whole:
            return _resolve_ref(root, whole.group(1), stack)
        prev = None
        while prev != node and "${" in node:
            prev = node
            node = _INTERP.sub(lambda m: str(_resolve_ref(root, m.group(1), stack)), node)
        return node
    return node


def resolve(cfg: ConfigNode) -> ConfigNode:
    """Resolve every ``${...}`` in place (OmegaConf.resolve equivalent)."""
    resolved = _resolve_node(cfg, cfg)
    return resolved if isinstance(resolved, ConfigNode) else ConfigNode(resolved)


# ----------------------------------------------------------------------------
# public entry point
# ----------------------------------------------------------------------------
DEFAULT_CONFIG_DIR = Path(__file__).resolve().parent.parent / "conf"


def compose(overrides: Iterable[str] = (), config_name: str = "base",
            config_dir: str | os.PathLike | None = None, do_resolve: bool = True,
            validate: bool = True) -> ConfigNode:
    """Compose ``conf/<config_name>.yaml`` + defaults + overrides → resolved tree."""
    cdir = Path(config_dir) if config_dir is not None else DEFAULT_CONFIG_DIR
    choices, values = split_overrides(cdir, list(overrides))
    used: dict[str, Any] = {}
    tree = _compose_file(cdir, config_name, choices, "", used)
    unknown = [k for k in choices if k not in used]
    if unknown:
        raise ConfigError(f"group override(s) {unknown} do not match any defaults entry; have {sorted(used)}")
    cfg = ConfigNode(tree)
    apply_overrides(cfg, values)
    if do_resolve:
        cfg = resolve(cfg)
    if validate:
        from photon_b200.config.schema import validate_config

        model = validate_config(cfg)
        # keys the schema knows but the YAML tree does not spell out (e.g. a reference user's own conf/ directory, which has no
        # ``photon.comm_stack.nvl`` / ``photon.topology`` / ``kernels``) are materialised with their defaults, so the dumped
        # config.yaml is complete and downstream code never has to guess
        _fill_defaults(cfg, model.model_dump(mode="json", exclude={"llm_config", "dataset", "eval_gauntlet_config", "icl_tasks_config"}))
    return cfg


def _fill_defaults(node: Any, defaults: dict[str, Any]) -> None:
    for k, v in defaults.items():
        if k not in node:
            node[k] = ConfigNode(v) if isinstance(v, dict) else v
        elif isinstance(v, dict) and isinstance(node[k], dict):
            _fill_defaults(node[k], v)
