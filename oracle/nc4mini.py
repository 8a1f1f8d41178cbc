"""TEST INFRASTRUCTURE -- minimal reader for the NetCDF-4 (= HDF5, superblock v2) field files that the
reference's tests ship (tests/data/{uinit,ufinal,eq}.nc).  NetCDF/HDF5 libraries are not installed in this
image, so this parses just enough of the HDF5 container: version-2 object headers, compact link messages,
contiguous little-endian float64 datasets, and the global attributes Nx,Ny,Nz,Lx,Lz,a,b (stored as dense
attributes in a fractal-heap direct block; located by their attribute-message header).

Follows the semantics of FlowField::readNetCDF (/root/reference/channelflow/flowfield.cpp:3605-3826):
variables Velocity_X/Y/Z have dims (Z, Y, X) on the dealiased I/O grid; attributes hold the full grid.
"""
import struct
import numpy as np


def _u(b, off, n):
    return int.from_bytes(b[off:off + n], "little")


def _parse_messages(b, start, end, out, ochk_list):
    off = start
    while off + 4 <= end:
        mtype = b[off]
        msize = _u(b, off + 1, 2)
        mflags = b[off + 3]
        off += 4
        if out["track_order"]:
            off += 2
        body = off
        if mtype == 0x10:  # continuation
            ochk_list.append((_u(b, body, 8), _u(b, body + 8, 8)))
        elif mtype == 0x06:  # link
            p = body
            ver, fl = b[p], b[p + 1]
            p += 2
            ltype = 0
            if fl & 0x08:
                ltype = b[p]; p += 1
            if fl & 0x04:
                p += 8
            if fl & 0x10:
                p += 1
            nlen_sz = 1 << (fl & 3)
            nlen = _u(b, p, nlen_sz); p += nlen_sz
            name = b[p:p + nlen].decode(); p += nlen
            if ltype == 0:
                out["links"][name] = _u(b, p, 8)
        elif mtype == 0x01:  # dataspace
            ver, rank, fl = b[body], b[body + 1], b[body + 2]
            p = body + (8 if ver == 1 else 4)
            out["dims"] = [_u(b, p + 8 * i, 8) for i in range(rank)]
        elif mtype == 0x08:  # layout
            ver, cls = b[body], b[body + 1]
            if ver in (3, 4) and cls == 1:
                out["addr"] = _u(b, body + 2, 8)
                out["size"] = _u(b, body + 10, 8)
            else:
                out["layout_unsupported"] = (ver, cls)
        off = body + msize
    return


def _parse_ohdr(b, addr):
    assert b[addr:addr + 4] == b"OHDR", "object header v2 expected"
    ver, fl = b[addr + 4], b[addr + 5]
    p = addr + 6
    if fl & 0x20:
        p += 16
    if fl & 0x10:
        p += 4
    csz = 1 << (fl & 3)
    chunk0 = _u(b, p, csz); p += csz
    out = {"links": {}, "track_order": bool(fl & 0x04)}
    ochk = []
    _parse_messages(b, p, p + chunk0, out, ochk)
    while ochk:
        a, ln = ochk.pop(0)
        assert b[a:a + 4] == b"OCHK"
        _parse_messages(b, a + 4, a + ln - 4, out, ochk)
    return out


def _find_attr(b, name):
    """Locate an attribute message (version 3) by name anywhere in the file and decode a scalar value."""
    nm = name.encode() + b"\0"
    start = 0
    while True:
        i = b.find(nm, start)
        if i < 0:
            raise KeyError(name)
        h = i - 9
        if h >= 0 and b[h] == 3 and _u(b, h + 2, 2) == len(nm):
            dts, dss = _u(b, h + 4, 2), _u(b, h + 6, 2)
            dt = i + len(nm)
            cls = b[dt] & 0x0F
            size = _u(b, dt + 4, 4)
            data = dt + dts + dss
            if cls == 0 and size == 4:
                return struct.unpack_from("<i", b, data)[0]
            if cls == 1 and size == 8:
                return struct.unpack_from("<d", b, data)[0]
        start = i + 1


def read_nc(path):
    """Returns (attrs, data) with attrs = dict(Nx,Ny,Nz,Lx,Lz,a,b) and data = float64 array [Nd][Nz_io][Ny][Nx_io]."""
    b = open(path, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 2, "HDF5 superblock v2 expected"
    so = b[9]
    assert so == 8
    root_addr = _u(b, 12 + 3 * so, so)
    root = _parse_ohdr(b, root_addr)
    comps = []
    for nm in ("Velocity_X", "Velocity_Y", "Velocity_Z"):
        if nm not in root["links"]:
            continue
        d = _parse_ohdr(b, root["links"][nm])
        assert "addr" in d, d
        n = int(np.prod(d["dims"]))
        assert d["size"] == 8 * n
        comps.append(np.frombuffer(b, dtype="<f8", count=n, offset=d["addr"]).reshape(d["dims"]).copy())
    attrs = {k: _find_attr(b, k) for k in ("Nx", "Ny", "Nz", "Lx", "Lz", "a", "b")}
    return attrs, np.stack(comps)


if __name__ == "__main__":
    import sys
    a, d = read_nc(sys.argv[1])
    print(a, d.shape, float(np.abs(d).max()))
