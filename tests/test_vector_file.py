"""Vector files (src/utils.rs:34-107, src/api.rs:149-186): concatenated bincode-2 `standard()` batches of
(id u64, values Vec<f32>, metadata u64).  The encoder/decoder below restate bincode's varint rule independently
of the library: u < 251 one byte; 251 + u16, 252 + u32, 253 + u64 little endian."""
import struct

import numpy as np
import pytest


def varint(v):
    if v < 251:
        return bytes([v])
    if v < 1 << 16:
        return b"\xfb" + struct.pack("<H", v)
    if v < 1 << 32:
        return b"\xfc" + struct.pack("<I", v)
    return b"\xfd" + struct.pack("<Q", v)


def encode_batch(recs):
    out = varint(len(recs))
    for i, v, m in recs:
        out += varint(i) + varint(len(v)) + np.asarray(v, "<f4").tobytes() + varint(m)
    return out


def read_varint(b, pos):
    t = b[pos]
    if t < 251:
        return t, pos + 1
    w = {251: 2, 252: 4, 253: 8}[t]
    return int.from_bytes(b[pos + 1:pos + 1 + w], "little"), pos + 1 + w


def decode_file(b):
    pos, recs = 0, []
    while pos < len(b):
        n, pos = read_varint(b, pos)
        for _ in range(n):
            i, pos = read_varint(b, pos)
            ln, pos = read_varint(b, pos)
            v = np.frombuffer(b, "<f4", ln, pos).copy()
            pos += 4 * ln
            m, pos = read_varint(b, pos)
            recs.append((i, v, m))
    return recs


@pytest.fixture
def ffi():
    from vector_indexer_py import _ffi
    return _ffi


def make_recs(n, d, seed=1):
    rng = np.random.default_rng(seed)
    ids = [int(x) for x in rng.integers(0, 2 ** 40, n)]
    ids[0], ids[1 % n] = 250, 251  # both sides of the one-byte varint boundary
    meta = [int(x) for x in rng.integers(0, 2 ** 63, n)]
    meta[0] = 0
    return [(ids[i], rng.standard_normal(d).astype(np.float32), meta[i]) for i in range(n)]


def test_reader_decodes_independently_encoded_batches(tmp_path, ffi):
    recs = make_recs(2500, 7)
    path = tmp_path / "v.bin"
    path.write_bytes(encode_batch(recs[:1000]) + encode_batch(recs[1000:2000]) + encode_batch(recs[2000:]))
    ids, data, meta = ffi.read_vector_file(str(path), 7)
    assert ids.tolist() == [r[0] for r in recs]
    assert meta.tolist() == [r[2] for r in recs]
    assert np.array_equal(data.view(np.uint32), np.stack([r[1] for r in recs]).view(np.uint32))


def test_reader_stops_silently_at_a_broken_batch(tmp_path, ffi):
    # utils.rs:98-104: `Err(_) => break` keeps the batches decoded so far
    recs = make_recs(30, 4)
    good = encode_batch(recs[:20])
    path = tmp_path / "v.bin"
    path.write_bytes(good + encode_batch(recs[20:])[:-5])
    ids, data, _ = ffi.read_vector_file(str(path), 4)
    assert ids.tolist() == [r[0] for r in recs[:20]] and data.shape == (20, 4)


def test_writer_produces_the_reference_framing(tmp_path, ffi):
    recs = make_recs(2300, 5, seed=3)
    path = tmp_path / "w.bin"
    ffi.write_vector_file(str(path), np.stack([r[1] for r in recs]), [r[0] for r in recs], [r[2] for r in recs], batch=1000)
    b = path.read_bytes()
    assert b == encode_batch(recs[:1000]) + encode_batch(recs[1000:2000]) + encode_batch(recs[2000:])
    back = decode_file(b)
    assert [r[0] for r in back] == [r[0] for r in recs]


def test_missing_file_is_an_error(tmp_path, ffi):
    with pytest.raises(ffi.VidxError):
        ffi.read_vector_file(str(tmp_path / "nope.bin"), 4)


@pytest.mark.gpu
def test_build_from_vector_file_equals_build_from_records(tmp_path, ffi):
    # api.rs:149-186 vs :115-146: same store, same seed -> same index
    rng = np.random.default_rng(5)
    xb = rng.standard_normal((6000, 24)).astype(np.float32)
    ids = np.arange(6000, dtype=np.uint64) * 3 + 7
    path = tmp_path / "v.bin"
    ffi.write_vector_file(str(path), xb, ids, np.full(6000, 1234, np.uint64))
    a = ffi.Index(24).build_from_vector_file(str(path))
    b = ffi.Index(24).build(xb, ext_ids=ids, timestamps=np.full(6000, 1234, np.uint64))
    assert a.nlist == b.nlist and np.array_equal(a.centroids().view(np.uint32), b.centroids().view(np.uint32))
    xq = rng.standard_normal((64, 24)).astype(np.float32)
    Da, Ia = a.search(xq, 10, 8)
    Db, Ib = b.search(xq, 10, 8)
    assert np.array_equal(Da.view(np.uint32), Db.view(np.uint32)) and np.array_equal(Ia, Ib)
    assert set(Ia.ravel().tolist()) <= set(ids.tolist())


@pytest.mark.gpu
def test_build_from_vector_file_errors(tmp_path, ffi):
    # api.rs:158-179
    empty = tmp_path / "e.bin"
    empty.write_bytes(b"")
    with pytest.raises(ffi.InvalidInput, match="no vectors in vector_file"):
        ffi.Index(8).build_from_vector_file(str(empty))
    bad = tmp_path / "b.bin"
    bad.write_bytes(encode_batch([(1, np.zeros(8, np.float32), 0), (2, np.zeros(5, np.float32), 0)]))
    with pytest.raises(ffi.InvalidInput, match="vector dimension mismatch at index 1: expected 8, got 5"):
        ffi.Index(8).build_from_vector_file(str(bad))
