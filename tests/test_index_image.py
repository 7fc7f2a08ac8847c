"""Index image (<prefix>.veb / <prefix>.vemb): bbq_index_save / bbq_index_load against the host-side reader.
Reference anchors: serializeVectorData / deserializeVectorData src/binaryQuantizationFormat.ts:483-560,
MetadataFormat / VectorDataFormat src/types.ts:78-113, FILE_EXTENSIONS src/constants.ts:52-57."""
import os
import struct

import numpy as np
import pytest

from tests.fixtures import gaussian


def _image():
    import importlib
    import bbq_b200  # noqa: F401  (registers the package under its importable name)
    return importlib.import_module("better_binary_quantization_b200.host.image")


@pytest.fixture(scope="module")
def bbq():
    import bbq_b200
    bbq_b200.build_library()
    return bbq_b200


def test_header_layout_matches_the_native_struct():
    img = _image()
    assert img.HEADER_BYTES == 144            # static_assert(sizeof(MetaHeader) == 144) in csrc/bbq_io.cuh
    assert img.row_bytes_for(1024) == 128 and img.row_bytes_for(768) == 96 and img.row_bytes_for(100) == 16
    lay = img.section_layout(1000, 128)
    assert [l[0] for l in lay] == list(img.SECTIONS)
    assert all(off % img.SECTION_ALIGN == 0 for _, off, _ in lay)
    assert lay[0][2] == 1000 * 16 and lay[1][2] == 8000 and lay[4][2] == 4000


def test_checksum_is_position_dependent_and_additive():
    img = _image()
    a = np.arange(64, dtype=np.uint32)
    b = a.copy()
    b[[3, 4]] = b[[4, 3]]
    assert img.section_checksum(a.tobytes()) != img.section_checksum(b.tobytes())
    assert img.section_checksum(b"") == 0
    # known answer: one zero word -> splitmix64(0)
    assert img.section_checksum(struct.pack("<I", 0)) == 0xE220A8397B1DCDAF


def test_reader_rejects_foreign_files(tmp_path):
    img = _image()
    p = str(tmp_path / "x")
    open(p + ".vemb", "wb").write(b"\0" * 200)
    with pytest.raises(img.ImageFormatError):
        img.read_metadata(p)
    open(p + ".vemb", "wb").write(b"BVEC")
    with pytest.raises(img.ImageFormatError):
        img.read_metadata(p)


@pytest.mark.gpu
@pytest.mark.parametrize("sim,n,dim", [("COSINE", 1000, 128), ("EUCLIDEAN", 3001, 100), ("MAXIMUM_INNER_PRODUCT", 777, 768)])
def test_save_load_round_trip(bbq, tmp_path, sim, n, dim):
    img = _image()
    from tests.test_gpu_parity import make_format
    fmt = make_format(bbq, sim)
    rows = gaussian(n, dim, seed=31)
    qv = fmt.quantizeVectors(rows)["quantizedVectors"]
    prefix = str(tmp_path / "index")
    fmt.saveIndex(qv, prefix)
    assert os.path.exists(prefix + ".veb") and os.path.exists(prefix + ".vemb")

    # host reader sees the reference's MetadataFormat / VectorDataFormat fields, checksums verified in numpy
    meta = img.read_metadata(prefix)
    packed, corr = qv.exportAll()
    assert meta["vectorCount"] == n and meta["dimensions"] == dim and meta["fieldNumber"] == 0
    assert meta["similarityFunction"] == sim and meta["indexBits"] == 1
    assert np.array_equal(meta["centroid"].view(np.uint32), np.asarray(qv.getCentroid(), np.float32).view(np.uint32))
    assert meta["centroidSquareMagnitude"] == qv.getCentroidDP()
    data = img.read_vector_data(prefix, meta, verify=True)
    assert np.array_equal(data["binaryValues"], packed)
    assert np.array_equal(data["lowerInterval"].view(np.uint64), corr[:, 0].copy().view(np.uint64))
    assert np.array_equal(data["upperInterval"].view(np.uint64), corr[:, 1].copy().view(np.uint64))
    assert np.array_equal(data["additionalCorrection"].view(np.uint64), corr[:, 2].copy().view(np.uint64))
    assert np.array_equal(data["quantizedComponentSum"], corr[:, 3].astype(np.uint32))

    # native load: same bytes on the device, same search results bit for bit
    fmt2 = make_format(bbq, sim)
    qv2 = fmt2.loadIndex(prefix)
    p2, c2 = qv2.exportAll()
    assert np.array_equal(p2, packed) and np.array_equal(c2.view(np.uint64), corr.view(np.uint64))
    assert qv2.size() == n and qv2.dimension() == dim and qv2.getCentroidDP() == qv.getCentroidDP()
    queries = gaussian(9, dim, seed=32)
    i1, s1 = fmt.searchBatch(queries, qv, 10)
    i2, s2 = fmt2.searchBatch(queries, qv2, 10)
    assert np.array_equal(i1, i2) and np.array_equal(s1.view(np.uint32), s2.view(np.uint32))


@pytest.mark.gpu
def test_load_rejects_corruption_and_mismatch(bbq, tmp_path):
    from tests.test_gpu_parity import make_format
    fmt = make_format(bbq, "COSINE")
    qv = fmt.quantizeVectors(gaussian(500, 128, seed=33))["quantizedVectors"]
    prefix = str(tmp_path / "index")
    fmt.saveIndex(qv, prefix)

    # another similarity function must not adopt the image (its corrective terms mean something else)
    with pytest.raises(bbq.BbqError) as e:
        make_format(bbq, "EUCLIDEAN").loadIndex(prefix)
    assert e.value.status == 12

    # one flipped bit in the codes, one in a corrective array
    blob = bytearray(open(prefix + ".veb", "rb").read())
    for pos in (17, 8192 + 5):
        bad = bytearray(blob)
        bad[pos] ^= 0x10
        open(prefix + ".veb", "wb").write(bad)
        with pytest.raises(bbq.BbqError) as e:
            fmt.loadIndex(prefix)
        assert e.value.status == 12 and "checksum" in str(e.value)

    # truncation
    open(prefix + ".veb", "wb").write(blob[:len(blob) // 2])
    with pytest.raises(bbq.BbqError) as e:
        fmt.loadIndex(prefix)
    assert e.value.status == 12
    open(prefix + ".veb", "wb").write(blob)
    assert fmt.loadIndex(prefix).size() == 500

    # missing file -> I/O error with the OS message
    with pytest.raises(bbq.BbqError) as e:
        fmt.loadIndex(str(tmp_path / "nope"))
    assert e.value.status == 11
    # unwritable target
    with pytest.raises(bbq.BbqError) as e:
        fmt.saveIndex(qv, str(tmp_path / "no_such_dir" / "index"))
    assert e.value.status == 11
