"""Shard file format (src/shards.rs:22-51, :68-177): layout pinned byte by byte from the
reference's #[repr(C)] structs, round trips as in tests/shards_tests.rs."""
import struct

import numpy as np
import pytest


def write_example(oracle, path, dim=3):
    cents = np.array([[1, 2, 3], [4, 5, 6]], np.float32)[:, :dim]
    lens = [2, 1]
    meta = np.array([[10, 100, 1000], [11, 101, 1001], [12, 102, 1002]], np.uint64)
    vecs = np.array([[0.1, 0.2, 0.3], [0.4, 0.5, 0.6], [0.7, 0.8, 0.9]], np.float32)[:, :dim]
    assert oracle.shard_write(path, 7, dim, [5, 2 ** 40 + 9], cents, lens, meta, vecs) == 0
    return cents, lens, meta, vecs


def test_byte_layout_dim3(oracle, tmp_path):
    """D=3: 12-byte vectors need 4 bytes of padding to keep VectorMeta 8-aligned
    (shards.rs:105-109); header is 40 bytes, index entries 32."""
    p = str(tmp_path / "shard_7.bin")
    cents, lens, meta, vecs = write_example(oracle, p)
    b = open(p, "rb").read()
    shard_id, version, dim, ncent, index_off, data_off = struct.unpack_from("<QQIIQQ", b, 0)
    assert (shard_id, version, dim, ncent, index_off, data_off) == (7, 1, 3, 2, 40, 40 + 64)
    cid0, n0, pad0, off0, size0 = struct.unpack_from("<QIIQQ", b, 40)
    cid1, n1, pad1, off1, size1 = struct.unpack_from("<QIIQQ", b, 72)
    rec = 24 + 12 + 4
    assert (cid0, n0, pad0, off0, size0) == (5, 2, 0, 104, 16 + 2 * rec)
    assert (cid1, n1, off1, size1) == (2 ** 40 + 9, 1, 104 + 16 + 2 * rec, 16 + rec)
    assert len(b) == off1 + size1
    assert struct.unpack_from("<3f", b, 104) == tuple(cents[0])
    assert struct.unpack_from("<QQQ", b, 104 + 16) == (10, 100, 1000)
    assert struct.unpack_from("<3f", b, 104 + 16 + 24) == tuple(vecs[0])
    assert struct.unpack_from("<QQQ", b, 104 + 16 + rec) == (11, 101, 1001)


def test_round_trip_preserves_metadata_exactly(oracle, tmp_path):
    # tests/shards_tests.rs:41-143, :358-408, :460-503
    p = str(tmp_path / "shard_7.bin")
    cents, lens, meta, vecs = write_example(oracle, p)
    r = oracle.shard_read(p, 7)
    assert r["dim"] == 3 and r["centroid_ids"].tolist() == [5, 2 ** 40 + 9] and r["lens"].tolist() == lens
    assert np.array_equal(r["centroid_vecs"], cents) and np.array_equal(r["meta"], meta) and np.array_equal(r["vecs"], vecs)


def test_empty_list_and_large_dim(oracle, tmp_path):
    # tests/shards_tests.rs:211-270
    p = str(tmp_path / "shard_0.bin")
    cents = np.random.default_rng(0).standard_normal((2, 512)).astype(np.float32)
    vecs = np.random.default_rng(1).standard_normal((3, 512)).astype(np.float32)
    meta = np.arange(9, dtype=np.uint64).reshape(3, 3)
    assert oracle.shard_write(p, 0, 512, [0, 1], cents, [0, 3], meta, vecs) == 0
    r = oracle.shard_read(p, 0)
    assert r["lens"].tolist() == [0, 3] and np.array_equal(r["vecs"], vecs)


def test_missing_file_wrong_id_and_corruption_are_errors(oracle, tmp_path):
    # tests/shards_tests.rs:541-630
    with pytest.raises(IOError):
        oracle.shard_read(str(tmp_path / "nope.bin"), 0)
    p = str(tmp_path / "shard_7.bin")
    write_example(oracle, p)
    with pytest.raises(IOError):
        oracle.shard_read(p, 8)  # shard id mismatch (shards.rs:223-231)
    with open(p, "r+b") as f:
        f.write(b"\xff" * 8)
    with pytest.raises(IOError):
        oracle.shard_read(p, 7)
