"""VTK restart files for the mirror: `load(sim, fname)` is `load!(sim::TwoPhaseSimulation, Val(:pvd); fname)` of the reference's
ReadVTK extension (ext/IntfAdvReadVTKExt.jl:29-50) -- read the LAST dataset of a ParaView collection (`.pvd`), check that its whole
extent matches the simulation's arrays, copy the point data `p`, `f` and `u` (components first in the file, last in the arrays) into
the device arrays, and reset the time step to the file's time.

The reference reads through ReadVTK.jl and the files are written by WaterLily's `vtkWriter` through WriteVTK.jl (third-party, not in
/root/reference; no sample file is held by the reference's tests, so this reader is written against the VTK XML ImageData format
itself): `<DataArray>` in `format="ascii"`, `"binary"` (inline base64) or `"appended"` (raw or base64), with or without the
`vtkZLibDataCompressor` block compression, `header_type` UInt32 / UInt64, little or big endian.  `write_vti` / `write_pvd` produce
the layout WriteVTK.jl emits by default (appended raw data, zlib blocks, UInt64 headers) so that runs of the mirror can be handed to
the reference (`load!`) and back.  Host-side format code: no GPU work besides the final copies.
"""
from __future__ import annotations

import base64
import os
import re
import struct
import xml.etree.ElementTree as ET
import zlib
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_VTK_DTYPES = {"Float32": "f4", "Float64": "f8", "Int8": "i1", "UInt8": "u1", "Int16": "i2", "UInt16": "u2", "Int32": "i4",
               "UInt32": "u4", "Int64": "i8", "UInt64": "u8"}
_NP_TO_VTK = {np.dtype(v).str[1:]: k for k, v in _VTK_DTYPES.items()}


class VTKFormatError(ValueError):
    pass


def read_pvd(fname: str) -> Tuple[List[float], List[str]]:
    """(timesteps, file names resolved against the collection's directory) of a ParaView `.pvd` collection."""
    root = ET.parse(fname).getroot()
    if root.tag != "VTKFile" or root.get("type") != "Collection":
        raise VTKFormatError(f"{fname}: not a VTK Collection file")
    base = os.path.dirname(os.path.abspath(fname))
    ts, files = [], []
    for ds in root.iter("DataSet"):
        ts.append(float(ds.get("timestep", "0")))
        files.append(os.path.join(base, ds.get("file")))
    if not files:
        raise VTKFormatError(f"{fname}: empty collection")
    return ts, files


def _split_appended(raw: bytes) -> Tuple[Optional[bytes], str, bytes]:
    """The XML parser cannot see raw appended data: cut the <AppendedData> payload out and return (payload, encoding, xml without it)."""
    m = re.search(rb"<AppendedData[^>]*encoding=\"(\w+)\"[^>]*>", raw)
    if not m:
        return None, "", raw
    start = raw.index(b"_", m.end()) + 1
    end = raw.rindex(b"</AppendedData>")
    payload = raw[start:end]
    xml = raw[:m.end()] + b"</AppendedData>" + raw[end + len(b"</AppendedData>"):]
    return payload, m.group(1).decode(), xml


def _decode_block(buf: bytes, pos: int, hdr: np.dtype, compressed: bool, b64: bool) -> bytes:
    """One data array starting at byte `pos` of `buf` (raw bytes, or base64 text when b64)."""
    hs = hdr.itemsize
    if not b64:
        if not compressed:
            n = int(np.frombuffer(buf, hdr, 1, pos)[0])
            return buf[pos + hs:pos + hs + n]
        nb, bs, last = (int(v) for v in np.frombuffer(buf, hdr, 3, pos))
        sizes = np.frombuffer(buf, hdr, nb, pos + 3 * hs).astype(np.int64)
        p = pos + (3 + nb) * hs
        out = []
        for s in sizes:
            out.append(zlib.decompress(buf[p:p + int(s)]))
            p += int(s)
        data = b"".join(out)
        expect = (nb - 1) * bs + (last if last else bs) if nb else 0
        if len(data) != expect:
            raise VTKFormatError("compressed block sizes do not add up")
        return data
    # base64: the header and the data are encoded as separate base64 streams (header first)
    text = buf[pos:]

    def b64take(t: bytes, nbytes: int) -> Tuple[bytes, int]:
        nchar = (nbytes + 2) // 3 * 4
        return base64.b64decode(t[:nchar])[:nbytes], nchar

    if not compressed:
        h, used = b64take(text, hs)
        n = int(np.frombuffer(h, hdr, 1)[0])
        # header and data may also be ONE stream (VTK >= 0.1 writers differ): try the joint form first
        joint = base64.b64decode(text[:(hs + n + 2) // 3 * 4])
        if len(joint) >= hs + n:
            return joint[hs:hs + n]
        d, _ = b64take(text[used:], n)
        return d
    h3, used = b64take(text, 3 * hs)
    nb, bs, last = (int(v) for v in np.frombuffer(h3, hdr, 3))
    hall, used = b64take(text, (3 + nb) * hs)
    sizes = np.frombuffer(hall, hdr, nb, 3 * hs).astype(np.int64)
    comp, _ = b64take(text[used:], int(sizes.sum()))
    out, p = [], 0
    for s in sizes:
        out.append(zlib.decompress(comp[p:p + int(s)]))
        p += int(s)
    return b"".join(out)


def read_vti(fname: str) -> Tuple[Tuple[int, ...], Dict[str, np.ndarray]]:
    """(whole extent as point counts per axis, point data) of a VTK XML ImageData file.  Arrays come back as numpy arrays of shape
    (ncomp, n1, n2, n3) in the file's component-first, x-fastest order (ReadVTK's `get_data_reshaped`), scalars as (n1, n2, n3)."""
    raw = open(fname, "rb").read()
    payload, enc, xml = _split_appended(raw)
    root = ET.fromstring(xml)
    if root.tag != "VTKFile" or root.get("type") != "ImageData":
        raise VTKFormatError(f"{fname}: not a VTK ImageData file")
    bo = "<" if root.get("byte_order", "LittleEndian") == "LittleEndian" else ">"
    hdr = np.dtype(bo + _VTK_DTYPES[root.get("header_type", "UInt32")])
    compressed = root.get("compressor") is not None
    if compressed and root.get("compressor") != "vtkZLibDataCompressor":
        raise VTKFormatError(f"{fname}: unsupported compressor {root.get('compressor')}")
    img = root.find("ImageData")
    ext = [int(v) for v in img.get("WholeExtent").split()]
    npts = tuple(ext[2 * i + 1] - ext[2 * i] + 1 for i in range(3))
    pd = img.find("Piece").find("PointData")
    out: Dict[str, np.ndarray] = {}
    for da in (pd.findall("DataArray") if pd is not None else []):
        dt = np.dtype(bo + _VTK_DTYPES[da.get("type")])
        nc = int(da.get("NumberOfComponents", "1"))
        fmt = da.get("format", "ascii")
        if fmt == "ascii":
            a = np.array(da.text.split(), dtype=dt)
        elif fmt == "binary":
            a = np.frombuffer(_decode_block((da.text or "").strip().encode(), 0, hdr, compressed, True), dt)
        elif fmt == "appended":
            if payload is None:
                raise VTKFormatError(f"{fname}: appended array without <AppendedData>")
            a = np.frombuffer(_decode_block(payload, int(da.get("offset", "0")), hdr, compressed, enc == "base64"), dt)
        else:
            raise VTKFormatError(f"{fname}: unknown DataArray format {fmt}")
        n = npts[0] * npts[1] * npts[2]
        if a.size != n * nc:
            raise VTKFormatError(f"{fname}: array {da.get('Name')} has {a.size} values, expected {n * nc}")
        a = a.astype(dt.newbyteorder("="))
        # file order: components fastest, then x, y, z
        out[da.get("Name")] = a.reshape((nc,) + npts, order="F") if nc > 1 else a.reshape(npts, order="F")
    return npts, out


def write_vti(fname: str, point_data: Dict[str, np.ndarray], compress: bool = True, block: int = 1 << 15) -> None:
    """Write point data as WriteVTK.jl does by default: appended raw data, UInt64 headers, zlib block compression.  Scalars are
    (n1, n2[, n3]) arrays, vectors (ncomp, n1, n2[, n3]) with the components first."""
    first = next(iter(point_data.values()))
    dims = first.shape if first.ndim <= 3 and not _is_vec(first, point_data) else first.shape[1:]
    npts = tuple(dims) + (1,) * (3 - len(dims))
    chunks, arrays, off = [], [], 0
    for name, a in point_data.items():
        vec = a.shape != tuple(dims)
        nc = a.shape[0] if vec else 1
        data = np.asfortranarray(a).tobytes(order="F")
        if compress:
            blocks = [data[i:i + block] for i in range(0, len(data), block)] or [b""]
            comp = [zlib.compress(b, 6) for b in blocks]
            last = len(blocks[-1]) if len(blocks[-1]) != block else 0
            head = struct.pack("<%dQ" % (3 + len(comp)), len(blocks), block, last, *[len(c) for c in comp])
            blob = head + b"".join(comp)
        else:
            blob = struct.pack("<Q", len(data)) + data
        arrays.append((name, _NP_TO_VTK[np.dtype(a.dtype).str[1:]], nc, off))
        chunks.append(blob)
        off += len(blob)
    ext = " ".join(f"0 {n - 1}" for n in npts)
    head = ['<?xml version="1.0" encoding="utf-8"?>',
            '<VTKFile type="ImageData" version="1.0" byte_order="LittleEndian" header_type="UInt64"' +
            (' compressor="vtkZLibDataCompressor">' if compress else '>'),
            f'  <ImageData WholeExtent="{ext}" Origin="0.0 0.0 0.0" Spacing="1.0 1.0 1.0">', f'    <Piece Extent="{ext}">', '      <PointData>']
    for name, ty, nc, o in arrays:
        head.append(f'        <DataArray type="{ty}" Name="{name}" NumberOfComponents="{nc}" format="appended" offset="{o}"/>')
    head += ['      </PointData>', '    </Piece>', '  </ImageData>', '  <AppendedData encoding="raw">', '_']
    with open(fname, "wb") as fh:
        fh.write("\n".join(head).encode())
        fh.write(b"".join(chunks))
        fh.write(b"\n  </AppendedData>\n</VTKFile>\n")


def _is_vec(a: np.ndarray, pd: Dict[str, np.ndarray]) -> bool:
    shapes = {v.shape for v in pd.values()}
    return any(len(s) == a.ndim - 1 and a.shape[1:] == s for s in shapes)


def write_pvd(fname: str, timesteps: Sequence[float], files: Sequence[str]) -> None:
    lines = ['<?xml version="1.0" encoding="utf-8"?>', '<VTKFile type="Collection" version="1.0" byte_order="LittleEndian">', '  <Collection>']
    for t, f in zip(timesteps, files):
        lines.append(f'    <DataSet timestep="{t!r}" part="0" file="{f}"/>')
    lines += ['  </Collection>', '</VTKFile>', '']
    open(fname, "w").write("\n".join(lines))


def load(sim, fname: str = "WaterLily.pvd") -> float:
    """load!(sim, Val(:pvd); fname)  (ext/IntfAdvReadVTKExt.jl:29-50): the last dataset of the collection overwrites sim.intf.f and
    sim.flow.u (and sim.flow.p when the simulation carries a pressure array); Δt[end] becomes the file's time in simulation units.
    Returns that time.  Raises when the whole extent does not match the simulation's arrays (the reference's @assert)."""
    import torch

    from . import api

    ts, files = read_pvd(fname)
    npts, pd = read_vti(files[-1])
    a, c = sim.flow, sim.intf
    D = a.D
    extent = [n for n in npts if n != 1]  # filter(!iszero, whole_extent[2:2:end]) .+ 1
    if extent != list(c.f.shape):
        raise ValueError("The dimensions of the simulation do not match the dimensions of the vtk file.")
    dev = c.f.device

    def put(dst: "torch.Tensor", src: np.ndarray):
        nd = np.float32 if dst.dtype == torch.float32 else np.float64
        if tuple(src.shape) != tuple(dst.shape):
            raise ValueError("The dimensions of the simulation do not match the dimensions of the vtk file.")
        dst.copy_(api.from_numpy(np.asfortranarray(src.astype(nd, copy=False)), device=dev))

    squeeze = lambda x: x.reshape(tuple(n for n in x.shape if n != 1), order="F")
    if "f" in pd:
        put(c.f, squeeze(pd["f"]))
    if "p" in pd and getattr(a, "p", None) is not None:
        put(a.p, squeeze(pd["p"]))
    if "u" not in pd:
        raise VTKFormatError(f"{files[-1]}: no point data 'u'")
    u = pd["u"]  # (ncomp, n1, n2, n3): components_last + squeeze, keeping the first D components (2-D files carry a zero third one)
    u = np.moveaxis(u, 0, -1)
    u = u.reshape(tuple(n for n in u.shape[:-1] if n != 1) + (u.shape[-1],), order="F")[..., :D]
    put(a.u, u)
    t = ts[-1] * sim.L / sim.U
    a.dt[-1] = t
    a.dt.append(api.MPCFL(a, c))  # the reference pushes WaterLily.CFL(a.flow); the two-phase limit is the one this path steps with
    return t


def save(sim, fname: str = "WaterLily.pvd", t: Optional[float] = None) -> str:
    """Counterpart for the mirror: append the current f and u (components first, three of them, as VTK wants) as a new `.vti` dataset of
    the collection `fname` (created if absent)."""
    from . import api

    a, c = sim.flow, sim.intf
    base = os.path.splitext(os.path.abspath(fname))[0]
    ts, files = (read_pvd(fname) if os.path.exists(fname) else ([], []))
    k = len(files)
    vti = f"{base}_{k:06d}.vti"
    f = api.to_numpy(c.f)
    u = np.moveaxis(api.to_numpy(a.u), -1, 0)
    if a.D == 2:
        u = np.concatenate([u, np.zeros((1,) + u.shape[1:], u.dtype)], 0)
    write_vti(vti, {"f": f, "u": np.asfortranarray(u)})
    tt = api.sim_time(sim) if t is None else t
    write_pvd(fname, ts + [tt], [os.path.basename(p) for p in files] + [os.path.basename(vti)])
    return vti
