"""Property tests (hypothesis) for the pieces whose correctness is combinatorial rather than numerical: flat layouts,
shard bounds, streaming aggregation, the override grammar, time strings, device splitting, the running-mean update."""
import math

import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from photon_b200.config.composer import ConfigNode, apply_overrides, parse_value
from photon_b200.server.fleet import split_devices
from photon_b200.strategy.aggregation import StreamingMean, aggregate_parameters, naive_weighted_mean
from photon_b200.train.timestamp import Time
from photon_b200.train.trainer import shard_bounds
from photon_b200.utils.core import chunks_idx
from photon_b200.utils.flat import FlatLayout

# the same examples on every box: a property test must not turn red on somebody else's random seed
settings.register_profile("repeatable", derandomize=True, deadline=None, database=None)
settings.load_profile("repeatable")

names = st.text(alphabet="abcdefghijklmnopqrstuvwxyz._0123456789", min_size=1, max_size=12)
shapes = st.lists(st.integers(1, 7), min_size=0, max_size=3).map(tuple)


@settings(max_examples=60, deadline=None)
@given(st.dictionaries(names, shapes, min_size=1, max_size=12), st.sampled_from([1, 4, 64, 256]))
def test_flat_layout_is_sorted_aligned_disjoint_and_roundtrips(named, align):
    lay = FlatLayout.build(named.items(), align=align, total_multiple=align)
    assert list(lay.names) == sorted(named) and lay.n_params == sum(int(np.prod(s)) if s else 1 for s in named.values())
    spans = sorted(zip(lay.offsets, lay.numels))
    assert all(o % align == 0 for o, _ in spans) and all(a + n <= b for (a, n), (b, _) in zip(spans, spans[1:]))
    assert spans[-1][0] + spans[-1][1] <= lay.total and lay.total % align == 0
    flat = torch.arange(lay.total, dtype=torch.float32)
    back = torch.full((lay.total,), -1.0)
    lay.from_ndarrays(back, lay.to_ndarrays(flat))
    for i in range(len(lay.names)):
        assert torch.equal(lay.view(back, i), lay.view(flat, i))
    stacked = lay.stacked(("", "m/", "v/"))
    assert stacked.total == 3 * lay.total and len(stacked.names) == 3 * len(lay.names)
    assert torch.equal(stacked.view(torch.arange(stacked.total, dtype=torch.float32), "v/" + lay.names[0]).reshape(-1)[:1],
                       torch.tensor([2.0 * lay.total + lay.offsets[0]]))


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 1 << 22), st.integers(1, 8), st.sampled_from([1, 4, 256]))
def test_shard_bounds_tile_the_index_space(total, world, align):
    b = shard_bounds(total, world, align)
    assert len(b) == world and b[0][0] == 0 and b[-1][1] == total
    assert all(lo <= hi for lo, hi in b) and all(x[1] == y[0] for x, y in zip(b, b[1:]))
    assert all(lo % align == 0 for lo, _ in b[1:] if lo < total)


@settings(max_examples=50, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 2 ** 31 - 1), st.floats(0.5, 1e4)), min_size=1, max_size=9))
def test_streaming_mean_matches_textbook_mean(clients):
    xs = [(torch.randn(257, generator=torch.Generator().manual_seed(seed)), w) for seed, w in clients]
    sm, acc, n = StreamingMean(), None, 0.0
    for x, w in xs:
        sm.add(x, w)
        acc, n = aggregate_parameters(acc, x, n, w)
    want = naive_weighted_mean(xs)
    assert torch.allclose(sm.result(), want, atol=1e-5) and torch.allclose(acc, want, atol=1e-5)
    assert math.isclose(sm.total, sum(w for _, w in xs), rel_tol=1e-9) and math.isclose(n, sm.total, rel_tol=1e-9)


@settings(max_examples=100, deadline=None)
@given(st.integers(0, 64), st.integers(1, 9))
def test_chunks_and_device_groups_partition_exactly(n, k):
    spans = list(chunks_idx(list(range(n)), k))
    assert len(spans) == k and spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
    if n >= k:
        groups = split_devices(n, k)
        assert sorted(d for g in groups for d in g) == list(range(n)) and max(map(len, groups)) - min(map(len, groups)) <= 1


scalars = st.one_of(st.integers(-10 ** 6, 10 ** 6), st.booleans(), st.none(), st.floats(allow_nan=False, allow_infinity=False, width=32),
                    st.text(alphabet="abcxyz_-/", min_size=1, max_size=8))


@settings(max_examples=100, deadline=None)
@given(st.lists(st.tuples(st.lists(st.sampled_from(["a", "b", "c", "d"]), min_size=1, max_size=3).map(".".join), scalars), min_size=1, max_size=6))
def test_override_grammar_add_set_delete_roundtrip(assignments):
    """``++k=v`` always lands, a later ``k=v`` overrides it, ``~k`` removes it, values keep their YAML type."""
    cfg = ConfigNode({})
    final = {}
    for key, val in assignments:
        text = "null" if val is None else ("true" if val is True else "false" if val is False else repr(val) if isinstance(val, float) else str(val))
        # a key cannot be both a leaf and a parent: skip assignments that would contradict an earlier one
        if any(k != key and (k.startswith(key + ".") or key.startswith(k + ".")) for k in final):
            continue
        apply_overrides(cfg, _ov("++" + key + "=" + text))
        final[key] = parse_value(text)
    for key, want in final.items():
        node = cfg
        for part in key.split("."):
            node = node[part]
        assert node == want or (isinstance(want, float) and math.isclose(node, want, rel_tol=1e-6))
    for key in list(final):
        apply_overrides(cfg, _ov("~" + key))
        node, ok = cfg, True
        for part in key.split("."):
            if part not in node:
                ok = False
                break
            node = node[part]
        assert not ok


def _ov(text):
    from photon_b200.config.composer import DEFAULT_CONFIG_DIR, split_overrides

    return split_overrides(DEFAULT_CONFIG_DIR, [text])[1]


@settings(max_examples=100, deadline=None)
@given(st.integers(0, 10 ** 6), st.sampled_from(["ba", "ep", "sp", "tok"]))
def test_time_strings_roundtrip(value, unit):
    t = Time.parse(f"{value}{unit}")
    assert t.value == value and t.unit == unit and Time.parse(str(t)) == t
    if unit == "ba":
        assert t.to_batches() == value


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 12), st.lists(st.integers(1, 4), min_size=1, max_size=5), st.integers(0, 2 ** 16))
def test_work_queue_dispatches_every_client_exactly_once(n_clients, speeds, seed):
    """Whatever the node speeds and reply order, each sampled client runs exactly once, never two at a time on a node."""
    import random

    from photon_b200.server.server_util import ClientScheduler

    rng = random.Random(seed)
    nodes = list(range(100, 100 + len(speeds)))
    speed = dict(zip(nodes, speeds))
    running: dict[int, list[int]] = {}
    dispatched: list[int] = []

    def dispatch(node, cid):
        assert node not in running, "node was handed a second client while busy"
        running[node] = [cid, speed[node] + rng.randint(0, 2)]
        dispatched.append(cid)

    def poll():
        done = []
        for node in list(running):
            running[node][1] -= 1
            if running[node][1] <= 0:
                done.append((node, running.pop(node)[0], None))
        rng.shuffle(done)
        return done

    clients = rng.sample(range(1000), n_clients)
    got = [cid for _, cid, _ in ClientScheduler(clients, nodes, dispatch, poll)]
    assert sorted(got) == sorted(clients) == sorted(dispatched) and not running
    assert dispatched[: min(len(nodes), n_clients)] == clients[: min(len(nodes), n_clients)]   # first wave in sampling order


arrays = st.lists(st.tuples(st.lists(st.integers(1, 5), min_size=0, max_size=3), st.sampled_from(["float32", "float16", "int32", "int64"])), min_size=1, max_size=6)


@settings(max_examples=40, deadline=None)
@given(arrays, st.integers(0, 2 ** 16))
def test_shm_metadata_codec_and_segment_roundtrip(specs, seed):
    """Any mix of shapes / dtypes survives metadata → literal string → metadata and a trip through one flat POSIX segment."""
    import uuid

    from photon_b200.shm.utils import ModelParametersMetadata, get_parameters_shm, set_parameters_shm, unlink_quietly

    rng = np.random.default_rng(seed)
    arrs = [(rng.standard_normal(tuple(shp)) * 100).astype(dt) for shp, dt in specs]
    meta = ModelParametersMetadata.from_ndarrays(arrs)
    again = ModelParametersMetadata.from_literal(meta.to_literal())
    assert again == meta and all(lo % 64 == 0 and hi - lo == a.nbytes for (lo, hi), a in zip(meta.array_bounds, arrs))
    assert all(a[1] <= b[0] for a, b in zip(meta.array_bounds, meta.array_bounds[1:]))
    name = f"pbt_prop_{uuid.uuid4().hex[:10]}"
    shm, _ = set_parameters_shm(name, arrs)
    try:
        h, views = get_parameters_shm(name, again, copy=True)
        h.close()
        assert all(v.dtype == a.dtype and v.shape == a.shape and np.array_equal(v, a) for v, a in zip(views, arrs))
    finally:
        shm.close()
        unlink_quietly(name)


@settings(max_examples=30, deadline=None)
@given(st.lists(st.integers(0, 30), min_size=0, max_size=8, unique=True), st.sampled_from([None, -1, -2, 0, 3, 7]))
def test_resume_round_resolution(rounds, want):
    """``photon.resume_round``: None → fresh start; negative → counted from the newest COMPLETE round; k → must be complete."""
    import tempfile

    from photon_b200.checkpoint import CheckpointStore

    with tempfile.TemporaryDirectory() as tmp:
        store = CheckpointStore(tmp, "b")
        keys = ["current_server_parameters"]
        for r in rounds:
            d = store.round_dir("u", r)
            d.mkdir(parents=True)
            (d / "current_server_parameters.npz").write_bytes(b"x")
            if r % 5 != 4:                       # rounds ≡ 4 (mod 5) are left incomplete: no state.bin
                (d / "state.bin").write_bytes(b"x")
        complete = sorted(r for r in rounds if r % 5 != 4)
        assert store.obtain_sorted_rounds("u", keys) == complete
        if want is None:
            assert store.interpret_resume_round("u", None, keys) is None
        elif want < 0:
            got = store.interpret_resume_round("u", want, keys)
            assert got == (complete[want] if len(complete) >= -want else None)
        elif want in complete:
            assert store.interpret_resume_round("u", want, keys) == want
        else:
            try:
                store.interpret_resume_round("u", want, keys)
                raise AssertionError("an incomplete / missing round must not resolve")
            except FileNotFoundError:
                pass


# ------------------------------------------------------------------------------------------------ full sharding plan
@settings(max_examples=40, deadline=None)
@given(st.integers(1, 6), st.integers(1, 8), st.lists(st.integers(1, 700), min_size=3, max_size=6), st.integers(1, 3000))
def test_zero3_unit_plan_owns_every_index_exactly_once(n_layers, world, block_sizes, emb):
    """Any block structure, any world size: units tile the flat space, slices are equal and 256-aligned, the shards of all ranks
    reassemble the full vector, and the packed buffer of 1-D tensors is rebuilt from pieces of the owners' shards."""
    from photon_b200.parallel.zero3 import UnitPlan

    named = {}
    for b in range(n_layers):
        for j, sz in enumerate(block_sizes):
            named[f"transformer.blocks.{b}.t{j}.weight"] = (sz, 2) if j % 2 == 0 else (sz,)
    named["transformer.norm_f.weight"] = (7,)
    named["transformer.wte.weight"] = (emb, 3)
    lay = FlatLayout.build(named.items())
    pl = UnitPlan(lay, n_layers, world)
    spans = sorted((u.lo, u.hi) for u in pl.units)
    assert spans[0][0] == 0 and spans[-1][1] == lay.total and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert all(u.per % 256 == 0 and u.per * world >= u.hi - u.lo and u.off % 256 == 0 for u in pl.units)
    full = torch.arange(lay.total, dtype=torch.float32)
    owners = torch.zeros(lay.total)
    shards = []
    for r in range(world):
        sh = pl.full_to_shard(full, r, torch.empty(pl.shard_len))
        shards.append(sh)
        for u in range(len(pl.units)):
            lo, hi = pl.slice_range(u, r)
            owners[lo:hi] += 1
    assert bool((owners == 1).all())
    rec = torch.full((lay.total,), -1.0)
    for r in range(world):
        pl.shard_to_full(shards[r], r, rec)
    assert torch.equal(rec, full)
    small = torch.full((pl.small_len,), -1.0)
    for r, so, do, k in pl.small_copies():
        assert 0 <= r < world and so + k <= pl.shard_len
        small[do: do + k] = shards[r][so: so + k]
    for name, i in pl.small_index.items():
        o = pl.small_offsets[name]
        assert torch.equal(small[o: o + lay.numels[i]], full[lay.offsets[i]: lay.offsets[i] + lay.numels[i]])


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 9), st.sampled_from(["put", "get", "pop"])), min_size=1, max_size=60), st.integers(1, 6))
def test_client_state_cache_never_exceeds_its_budget(ops_, cap):
    """Whatever the access pattern: at most ``cap`` entries survive, a hit returns what was put last, and the survivors are the most
    recently used ones."""
    from photon_b200.federation import ClientStateCache

    one = 2 * 64 * 4
    c = ClientStateCache(device_bytes=0, host_bytes=cap * one)
    model: dict[int, int] = {}
    order: list[int] = []           # least recently used first
    for cid, op in ops_:
        if op == "put":
            stamp = len(order) + 1000 * cid
            c.put(cid, torch.full((64,), float(stamp)), torch.zeros(64), stamp)
            model[cid] = stamp
            order = [x for x in order if x != cid] + [cid]
            while len(order) > cap:
                model.pop(order.pop(0))
        elif op == "get" and cid in model:
            m, _, step = c.get(cid)
            assert step == model[cid] and float(m[0]) == float(model[cid])
            order = [x for x in order if x != cid] + [cid]
        elif op == "pop":
            got = c.pop(cid)
            assert (got is None) == (cid not in model)
            model.pop(cid, None)
            order = [x for x in order if x != cid]
        assert len(c) == len(model) <= cap and all(k in c for k in model)


@settings(max_examples=80, deadline=None)
@given(st.text(alphabet="abcXYZ019 +$&=?#%/._-~é", min_size=1, max_size=30), st.dictionaries(st.sampled_from(["prefix", "list-type", "uploads", "partNumber", "k y"]),
                                                                                           st.text(alphabet="ab 1/+=&", max_size=8), max_size=3))
def test_sigv4_signature_is_stable_under_odd_keys_and_queries(key, query):
    """The signed path / query are exactly what goes on the wire: an independent re-derivation (the test endpoint's) agrees for any key."""
    import datetime as dt
    import hashlib
    import hmac
    import urllib.parse

    from photon_b200.utils.objstore import _EMPTY_SHA, _quote, sigv4_headers

    now = dt.datetime(2024, 2, 29, 12, 0, 0, tzinfo=dt.timezone.utc)
    path = "/bkt/" + _quote(key.lstrip("/") or "k", safe="-_.~/")
    h = sigv4_headers("GET", "host:9000", path, query, {}, _EMPTY_SHA, access_key="AK", secret_key="s/e+c", region="r1", now=now)
    sig = h["authorization"].rsplit("Signature=", 1)[1]
    url_query = "&".join(f"{_quote(k)}={_quote(v)}" for k, v in sorted(query.items()))
    q = sorted(urllib.parse.parse_qsl(url_query, keep_blank_values=True))
    cq = "&".join(f"{urllib.parse.quote(k, safe='-_.~')}={urllib.parse.quote(v, safe='-_.~')}" for k, v in q)
    signed = "host;x-amz-content-sha256;x-amz-date"
    canonical = "\n".join(["GET", path, cq, f"host:host:9000\nx-amz-content-sha256:{_EMPTY_SHA}\nx-amz-date:20240229T120000Z\n", signed, _EMPTY_SHA])
    to_sign = "\n".join(["AWS4-HMAC-SHA256", "20240229T120000Z", "20240229/r1/s3/aws4_request", hashlib.sha256(canonical.encode()).hexdigest()])
    k = ("AWS4" + "s/e+c").encode()
    for part in ("20240229", "r1", "s3", "aws4_request"):
        k = hmac.new(k, part.encode(), hashlib.sha256).digest()
    assert hmac.new(k, to_sign.encode(), hashlib.sha256).hexdigest() == sig
    assert urllib.parse.unquote(path) == "/bkt/" + (key.lstrip("/") or "k")          # the key survives the encoding round trip
