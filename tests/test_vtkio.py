"""VTK restart format code (ext/IntfAdvReadVTKExt.jl:29-50): the reader against files written in every DataArray
encoding the VTK XML ImageData format allows, and the writer / reader round trip.  CPU only (the device copy of `load` is covered by
tests/test_gpu_post.py)."""
import base64
import struct
import zlib

import numpy as np
import pytest

from interfaceadvection.jl_b200 import vtkio


def _fields(shape, T, seed=0):
    rng = np.random.default_rng(seed)
    f = np.asfortranarray(rng.uniform(0, 1, shape).astype(T))
    u = np.asfortranarray(rng.standard_normal((3,) + shape).astype(T))
    return f, u


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(10, 7), (6, 5, 4)])
@pytest.mark.parametrize("compress", [True, False])
def test_write_read_round_trip(tmp_path, T, shape, compress):
    f, u = _fields(shape, T)
    fn = str(tmp_path / "a.vti")
    vtkio.write_vti(fn, {"f": f, "u": u}, compress=compress, block=256)  # small blocks: several zlib blocks per array
    npts, pd = vtkio.read_vti(fn)
    full = shape + (1,) * (3 - len(shape))
    assert npts == full
    assert pd["f"].dtype == np.dtype(T) and np.array_equal(pd["f"].reshape(shape, order="F"), f)
    assert np.array_equal(pd["u"].reshape((3,) + shape, order="F"), u)


def _xml(body, attrs="", appended=b""):
    return (b'<?xml version="1.0"?>\n<VTKFile type="ImageData" version="1.0" byte_order="LittleEndian" ' + attrs.encode() + b'>\n'
            b'<ImageData WholeExtent="0 2 0 1 0 0" Origin="0 0 0" Spacing="1 1 1"><Piece Extent="0 2 0 1 0 0"><PointData>\n' + body +
            b'\n</PointData></Piece></ImageData>\n' + appended + b'</VTKFile>\n')


def test_reader_handles_every_encoding(tmp_path):
    vals = np.arange(6, dtype=np.float32) * 0.5 - 1
    raw = vals.tobytes()
    want = vals.reshape((3, 2, 1), order="F")
    cases = {}
    cases["ascii"] = _xml(b'<DataArray type="Float32" Name="f" format="ascii">' + " ".join(map(str, vals)).encode() + b'</DataArray>')
    cases["binary32"] = _xml(b'<DataArray type="Float32" Name="f" format="binary">' + base64.b64encode(struct.pack("<I", len(raw)) + raw) +
                             b'</DataArray>', 'header_type="UInt32"')
    comp = zlib.compress(raw)
    head = struct.pack("<4Q", 1, len(raw), 0, len(comp))
    cases["binary64z"] = _xml(b'<DataArray type="Float32" Name="f" format="binary">' + base64.b64encode(head) + base64.b64encode(comp) +
                              b'</DataArray>', 'header_type="UInt64" compressor="vtkZLibDataCompressor"')
    cases["appended_raw"] = _xml(b'<DataArray type="Float32" Name="f" format="appended" offset="0"/>', 'header_type="UInt64"',
                                 b'<AppendedData encoding="raw">\n_' + struct.pack("<Q", len(raw)) + raw + b'\n</AppendedData>\n')
    cases["appended_b64"] = _xml(b'<DataArray type="Float32" Name="f" format="appended" offset="0"/>', 'header_type="UInt32"',
                                 b'<AppendedData encoding="base64">\n_' + base64.b64encode(struct.pack("<I", len(raw)) + raw) +
                                 b'\n</AppendedData>\n')
    for name, data in cases.items():
        fn = tmp_path / (name + ".vti")
        fn.write_bytes(data)
        npts, pd = vtkio.read_vti(str(fn))
        assert npts == (3, 2, 1), name
        assert np.array_equal(pd["f"], want), name


def test_big_endian_and_errors(tmp_path):
    vals = np.arange(6, dtype=">f8")
    raw = vals.tobytes()
    data = _xml(b'<DataArray type="Float64" Name="f" format="appended" offset="0"/>', 'header_type="UInt32"',
                b'<AppendedData encoding="raw">\n_' + struct.pack(">I", len(raw)) + raw + b'\n</AppendedData>\n').replace(b"LittleEndian", b"BigEndian")
    fn = tmp_path / "be.vti"
    fn.write_bytes(data)
    _, pd = vtkio.read_vti(str(fn))
    assert np.array_equal(pd["f"].ravel(order="F"), np.arange(6.0))
    bad = tmp_path / "bad.vti"
    bad.write_bytes(_xml(b'<DataArray type="Float32" Name="f" format="ascii">1 2 3</DataArray>'))
    with pytest.raises(vtkio.VTKFormatError):
        vtkio.read_vti(str(bad))
    notvti = tmp_path / "c.vti"
    notvti.write_text('<?xml version="1.0"?><VTKFile type="PolyData"></VTKFile>')
    with pytest.raises(vtkio.VTKFormatError):
        vtkio.read_vti(str(notvti))


def test_pvd_collection(tmp_path):
    fn = str(tmp_path / "WaterLily.pvd")
    vtkio.write_pvd(fn, [0.0, 0.5, 1.25], ["a_0.vti", "a_1.vti", "a_2.vti"])
    ts, files = vtkio.read_pvd(fn)
    assert ts == [0.0, 0.5, 1.25]
    assert files[-1] == str(tmp_path / "a_2.vti")
