import json

import numpy as np
import pytest
import yaml

from photon_b200.data.shards import ShardReader
from photon_b200.dataset.constants import DATASETS_CONSTANTS
from photon_b200.dataset.convert_dataset_hf import main as convert_main
from photon_b200.dataset.samples_generators import generate_samples_from_shards
from photon_b200.dataset.stream_partitioner import partition_stream_list, partition_streams
from photon_b200.dataset.utils import ByteTokenizer, concat_tokens


def test_constants_table():
    from photon_b200.dataset.constants import ConcatMode, resolve_split
    from photon_b200.dataset.constants import mc4

    assert len(DATASETS_CONSTANTS) == 13 and DATASETS_CONSTANTS["c4_en"].splits["train_small"].truncated_samples == 100_000
    # the reference's table shape (ref: photon/dataset/constants/mc4.py): six English entries, the two full splits elsewhere;
    # the full validation set is keyed "validation" and lands in the folder "val"
    en, it = DATASETS_CONSTANTS["c4_en"], DATASETS_CONSTANTS["c4_it"]
    assert list(en.splits) == ["train", "train_small", "validation", "val_small", "val_xsmall", "val_xxsmall"]
    assert [(s.split, s.folder_split, s.truncated_samples) for s in en][2:] == [("validation", "val", None), ("validation", "val_small", 10_000),
                                                                             ("validation", "val_xsmall", 3_000), ("validation", "val_xxsmall", 100)]
    assert list(it.splits) == ["train", "validation"] and it.splits["validation"].name == "it"
    assert resolve_split("c4_it", "val") is it.splits["validation"] and mc4.c4_hi_constants is DATASETS_CONSTANTS["c4_hi"]
    assert mc4.SERBIAN_CONSTANT == "sr" and ConcatMode("CONCAT_TOKENS") is ConcatMode.CONCAT_TOKENS
    import pytest
    with pytest.raises(KeyError, match="has no split"):
        resolve_split("c4_it", "val_xxsmall")


def test_concat_tokens_packs_with_eos():
    tok = ByteTokenizer()
    samples = list(concat_tokens(["ab", "cde", "f" * 10], tok, max_length=8))
    flat = np.concatenate(samples)
    assert all(s.shape == (8,) and s.dtype == np.int32 for s in samples)
    assert flat[2] == 0 and flat[6] == 0                     # eos after "ab" and after "cde"


def test_converter_contiguous_partitions_and_unigram(tmp_path):
    out = convert_main(["--source", "synthetic://200", "--out_root", str(tmp_path), "--num_clients", "4", "--concat_tokens", "32",
                        "--splits", "train_small", "--compression", "none"])
    counts = out["train_small"]
    assert len(counts) == 4 and max(counts) - min(counts) <= 3
    r = ShardReader(tmp_path / "c4" / "en" / "client_1" / "train_small")
    assert len(r) == counts[1] and r[0].shape == (32,)
    freq = json.loads((tmp_path / "c4" / "en" / "client_1" / "train_small" / "1_gram.json").read_text())
    assert sum(freq.values()) == counts[1] * 32
    assert (tmp_path / "tokenizer" / "tokenizer_config.json").exists()
    re64 = list(generate_samples_from_shards(tmp_path / "c4" / "en" / "client_1" / "train_small", new_seq_len=64))
    assert len(re64) == counts[1] // 2 and re64[0]["tokens"].shape == (64,)


def test_converter_reference_flags_single_client_and_mirror(tmp_path):
    """``--client k`` writes exactly the partition a full conversion gives client k; ``--remote_path`` mirrors the tree;
    ``--compression zstd`` / ``--num_workers`` / ``--pad_text`` are accepted like in the reference CLI."""
    common = ["--source", "synthetic://200", "--num_clients", "4", "--concat_tokens", "32", "--splits", "train_small", "--bos_text", "", "--num_workers", "2",
              "--pad_text", "<s>", "--tokenizer_kwargs", "{}"]
    full = convert_main(common + ["--out_root", str(tmp_path / "full"), "--compression", "zstd"])["train_small"]
    one = convert_main(common + ["--out_root", str(tmp_path / "one"), "--client", "2", "--compression", "none", "--remote_path", str(tmp_path / "mirror")])["train_small"]
    assert one[2] == full[2] and one[0] == one[1] == one[3] == 0
    a = ShardReader(tmp_path / "full" / "c4" / "en" / "client_2" / "train_small")
    b = ShardReader(tmp_path / "mirror" / "c4" / "en" / "client_2" / "train_small")
    assert len(a) == len(b) and all((a[i] == b[i]).all() for i in range(len(a)))
    assert not (tmp_path / "one" / "c4" / "en" / "client_0" / "train_small" / "index.json").exists()
    wrapped = convert_main(common + ["--out_root", str(tmp_path / "nw"), "--no_wrap"])["train_small"]
    assert sum(wrapped) <= sum(full)          # tails of documents are dropped instead of carried over


def test_stream_partitioner(tmp_path):
    entries = [{"client_streams": {f"stream_{i}": {"local": f"c8/en/client_{i}"}}} for i in range(8)]
    two = partition_stream_list(entries, 2)
    assert [len(e["client_streams"]) for e in two] == [4, 4] and two[1]["client_streams"]["stream_0"]["local"] == "c8/en/client_4"
    (tmp_path / "in.yaml").write_text(yaml.safe_dump(entries))
    partition_streams(tmp_path / "in.yaml", tmp_path / "out.yaml", 4)
    assert len(yaml.safe_load((tmp_path / "out.yaml").read_text())) == 4


@pytest.mark.parametrize("compression", ["zstd", "none"])
def test_converter_writes_mds_directories_the_loader_reads_back(tmp_path, compression):
    """``--format mds``: the converter writes mosaicml-streaming's layout (index v2, shard header / offsets / config / samples, the
    two-byte dynamic-shape header of ``ndarray:int32``); the MDS reader and the streaming loader read the same tokens as from the
    default shard format."""
    import json

    from photon_b200.data.shards import MDSReader, MDSWriter, open_shard_dir

    args = ["--dataset", "c4_en", "--splits", "train_small", "--source", "synthetic://60", "--num_clients", "2", "--concat_tokens", "300",
            "--tokenizer", "byte", "--compression", compression]
    a = convert_main(args + ["--out_root", str(tmp_path / "own")])
    b = convert_main(args + ["--out_root", str(tmp_path / "mds"), "--format", "mds"])
    assert a == b and a["train_small"][0] > 0
    for c in (0, 1):
        own = open_shard_dir(tmp_path / "own" / "c2" / "en" / f"client_{c}" / "train_small")
        mds = open_shard_dir(tmp_path / "mds" / "c2" / "en" / f"client_{c}" / "train_small")
        assert isinstance(mds, MDSReader) and len(mds) == len(own) == a["train_small"][c] and mds.seq_len == 300
        assert all(np.array_equal(mds[i], own[i]) for i in range(len(own)))
    d = tmp_path / "mds" / "c2" / "en" / "client_0" / "train_small"
    idx = json.loads((d / "index.json").read_text())
    sh = idx["shards"][0]
    assert idx["version"] == 2 and sh["format"] == "mds" and sh["column_encodings"] == ["ndarray:int32"] and sh["column_sizes"] == [None]
    assert (sh["compression"], sh["zip_data"] is None) == (("zstd", False) if compression == "zstd" else (None, True))
    # byte-level: 300 tokens need a uint16 dim -> header 01 02 2c 01, then 1200 data bytes
    w = MDSWriter(tmp_path / "one", seq_len=300, compression=None)
    w.write(np.arange(300, dtype=np.int32))
    w.finish()
    raw = (tmp_path / "one" / "shard.00000.mds").read_bytes()
    n, first, end = np.frombuffer(raw, np.uint32, 3)
    assert n == 1 and end == len(raw) and raw[first: first + 8] == np.uint32(1204).tobytes() + bytes([1, 2, 0x2C, 0x01])
    # several shards when the size limit is small; sample order survives
    w = MDSWriter(tmp_path / "many", seq_len=16, compression=None, size_limit=400)
    rows = [np.full(16, i, np.int32) for i in range(20)]
    w.write_many(rows)
    assert len(w.finish()["shards"]) > 1
    r = MDSReader(tmp_path / "many")
    assert len(r) == 20 and all(int(r[i][0]) == i for i in range(20))


def test_tokenizer_fallback_is_loud_and_strict_mode_refuses_it(monkeypatch, capsys):
    from photon_b200.dataset.utils import ByteTokenizer, build_tokenizer

    monkeypatch.delenv("PHOTON_STRICT_DATA", raising=False)
    assert isinstance(build_tokenizer("byte"), ByteTokenizer) and capsys.readouterr().err == ""       # asked for: silent
    tok = build_tokenizer("no-such-org/no-such-tokenizer")
    assert isinstance(tok, ByteTokenizer) and "BYTE-LEVEL fall-back" in capsys.readouterr().err
    monkeypatch.setenv("PHOTON_STRICT_DATA", "1")
    with pytest.raises(RuntimeError, match="not available locally"):
        build_tokenizer("no-such-org/no-such-tokenizer")
