// Reader for the NetCDF-4 field files Channelflow writes (FlowField::writeNetCDF / readNetCDF, reference
// channelflow/flowfield.cpp:3225-3826) without the NetCDF / HDF5 libraries (neither is in this image).  A NetCDF-4 file is
// an HDF5 container; the files of FlowField::writeNetCDF use superblock version 2, version-2 object headers with compact
// link messages, and contiguous little-endian float64 datasets Velocity_X/Y/Z (or Component_i) of shape (Z, Y, X) on the
// I/O grid, plus the scalar global attributes Nx, Ny, Nz, Lx, Lz, a, b of the full grid.  This parses exactly that subset
// and fails loudly on anything else (chunked or compressed layouts, old-style groups).  Writing NetCDF is not carried:
// FlowField::save writes the reference's binary .ff format, which stock Channelflow reads.
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "channelflow/flowfield.h"

namespace chflow {

namespace {

struct Hdf5 {
    std::vector<unsigned char> b;
    uint64_t le(size_t off, int n) const {
        if (off + n > b.size()) cferror("NetCDF reader: truncated file");
        uint64_t v = 0;
        for (int i = n - 1; i >= 0; --i) v = (v << 8) | b[off + i];
        return v;
    }
};

struct ObjHeader {
    std::map<std::string, uint64_t> links;  // hard links name -> object header address
    std::vector<uint64_t> dims;
    uint64_t addr = 0, size = 0;
    bool contiguous = false, track_order = false;
};

// header messages of one chunk: link (0x06), dataspace (0x01), layout (0x08), continuation (0x10)
void parse_messages(const Hdf5& f, size_t start, size_t end, ObjHeader& o, std::vector<std::pair<uint64_t, uint64_t>>& more) {
    size_t off = start;
    while (off + 4 <= end) {
        const int type = f.b[off];
        const size_t msize = (size_t)f.le(off + 1, 2);
        off += 4;
        if (o.track_order) off += 2;
        const size_t body = off;
        if (type == 0x10) more.emplace_back(f.le(body, 8), f.le(body + 8, 8));
        else if (type == 0x06) {
            size_t p = body;
            const int fl = f.b[p + 1];
            p += 2;
            int ltype = 0;
            if (fl & 0x08) ltype = f.b[p++];
            if (fl & 0x04) p += 8;
            if (fl & 0x10) p += 1;
            const int nsz = 1 << (fl & 3);
            const size_t nlen = (size_t)f.le(p, nsz);
            p += nsz;
            const std::string name((const char*)&f.b[p], nlen);
            p += nlen;
            if (ltype == 0) o.links[name] = f.le(p, 8);
        } else if (type == 0x01) {
            const int ver = f.b[body], rank = f.b[body + 1];
            const size_t p = body + (ver == 1 ? 8 : 4);
            o.dims.clear();
            for (int i = 0; i < rank; ++i) o.dims.push_back(f.le(p + 8 * i, 8));
        } else if (type == 0x08) {
            const int ver = f.b[body], cls = f.b[body + 1];
            if ((ver == 3 || ver == 4) && cls == 1) {
                o.addr = f.le(body + 2, 8);
                o.size = f.le(body + 10, 8);
                o.contiguous = true;
            }
        }
        off = body + msize;
    }
}

ObjHeader parse_object(const Hdf5& f, uint64_t addr) {
    if (addr + 6 > f.b.size() || memcmp(&f.b[addr], "OHDR", 4) != 0) cferror("NetCDF reader: version-2 object header expected");
    const int fl = f.b[addr + 5];
    size_t p = addr + 6;
    if (fl & 0x20) p += 16;
    if (fl & 0x10) p += 4;
    const int csz = 1 << (fl & 3);
    const size_t chunk0 = (size_t)f.le(p, csz);
    p += csz;
    ObjHeader o;
    o.track_order = (fl & 0x04) != 0;
    std::vector<std::pair<uint64_t, uint64_t>> more;
    parse_messages(f, p, p + chunk0, o, more);
    for (size_t k = 0; k < more.size(); ++k) {
        const uint64_t a = more[k].first, len = more[k].second;
        if (a + 4 > f.b.size() || memcmp(&f.b[a], "OCHK", 4) != 0) cferror("NetCDF reader: bad continuation block");
        parse_messages(f, a + 4, a + len - 4, o, more);
    }
    return o;
}

// scalar attribute (version-3 attribute message) by name: int32 or float64
double find_attr(const Hdf5& f, const std::string& name) {
    const std::string nm = name + std::string(1, '\0');
    for (size_t i = 9; i + nm.size() < f.b.size(); ++i) {
        if (memcmp(&f.b[i], nm.data(), nm.size()) != 0) continue;
        const size_t h = i - 9;
        if (f.b[h] != 3 || f.le(h + 2, 2) != nm.size()) continue;
        const size_t dts = (size_t)f.le(h + 4, 2), dss = (size_t)f.le(h + 6, 2);
        const size_t dt = i + nm.size();
        const int cls = f.b[dt] & 0x0F;
        const uint64_t size = f.le(dt + 4, 4);
        const size_t data = dt + dts + dss;
        if (cls == 0 && size == 4) { int32_t v; memcpy(&v, &f.b[data], 4); return v; }
        if (cls == 1 && size == 8) { double v; memcpy(&v, &f.b[data], 8); return v; }
    }
    cferror("NetCDF reader: attribute " + name + " not found");
    return 0;
}

}  // namespace

// Returns false if the file does not exist; fills u (resized, spectral, padded as the file says) otherwise.
bool read_netcdf_field(const std::string& filename, FlowField& u, CfMPI* cfmpi) {
    std::ifstream is(filename.c_str(), std::ios::in | std::ios::binary);
    if (!is.good()) return false;
    Hdf5 f;
    f.b.assign(std::istreambuf_iterator<char>(is), std::istreambuf_iterator<char>());
    static const unsigned char magic[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (f.b.size() < 64 || memcmp(f.b.data(), magic, 8) != 0 || f.b[8] != 2 || f.b[9] != 8)
        cferror("NetCDF reader: " + filename + " is not a NetCDF-4 file with an HDF5 version-2 superblock");
    const ObjHeader root = parse_object(f, f.le(12 + 3 * 8, 8));
    std::vector<ObjHeader> comps;
    static const char* vel[3] = {"Velocity_X", "Velocity_Y", "Velocity_Z"};
    for (const char* nm : vel)
        if (root.links.count(nm)) comps.push_back(parse_object(f, root.links.at(nm)));
    if (comps.empty())
        for (int i = 0; root.links.count("Component_" + std::to_string(i)); ++i) comps.push_back(parse_object(f, root.links.at("Component_" + std::to_string(i))));
    if (comps.empty()) cferror("NetCDF reader: no Velocity_X/Y/Z or Component_i variables in " + filename);
    const int Nx = (int)find_attr(f, "Nx"), Ny = (int)find_attr(f, "Ny"), Nz = (int)find_attr(f, "Nz");
    const Real Lx = find_attr(f, "Lx"), Lz = find_attr(f, "Lz"), a = find_attr(f, "a"), b = find_attr(f, "b");
    const int Nd = (int)comps.size();
    for (const ObjHeader& c : comps)
        if (!c.contiguous || c.dims.size() != 3 || c.size != 8 * c.dims[0] * c.dims[1] * c.dims[2] || c.dims != comps[0].dims ||
            c.addr + c.size > f.b.size())
            cferror("NetCDF reader: only contiguous float64 variables of shape (Z, Y, X) are supported: " + filename);
    const int Nz_io = (int)comps[0].dims[0], Ny_io = (int)comps[0].dims[1], Nx_io = (int)comps[0].dims[2];
    const bool full = Nx_io == Nx && Ny_io == Ny && Nz_io == Nz;
    if (!full && !(Nx_io == 2 * (Nx / 3) && Ny_io == Ny && Nz_io == 2 * (Nz / 3)))
        cferror("NetCDF reader: conflict between the file's dimensions and its grid attributes: " + filename);
    // physical values on the I/O grid, var[nx + Nx_io (ny + Ny nz)] (flowfield.cpp:3786-3790)
    FlowField g(Nx_io, Ny, Nz_io, Nd, Lx, Lz, a, b, cfmpi, Physical, Physical);
    for (int i = 0; i < Nd; ++i) {
        const unsigned char* base = &f.b[comps[i].addr];
        for (int nz = 0; nz < Nz_io; ++nz)
            for (int ny = 0; ny < Ny; ++ny)
                for (int nx = 0; nx < Nx_io; ++nx) {
                    double v;
                    memcpy(&v, base + 8 * ((size_t)nx + (size_t)Nx_io * (ny + (size_t)Ny * nz)), 8);
                    g(nx, ny, nz, i) = v;
                }
    }
    if (full) {
        g.makeSpectral();
        g.setPadded(false);
        u = g;
        return true;
    }
    // de-aliased I/O grid: transform in x,z there and place its modes into the full grid (addPaddedModes, flowfield.cpp:2991-3190:
    // rows mx <= Nx_io/2 stay, the negative-kx rows move to the end, every kz of the small grid is kept), then finish in y
    g.makeSpectral_xz();
    u = FlowField(Nx, Ny, Nz, Nd, Lx, Lz, a, b, cfmpi, Spectral, Physical);
    const int Mz_io = Nz_io / 2 + 1;
    for (int i = 0; i < Nd; ++i)
        for (int ny = 0; ny < Ny; ++ny)
            for (int mx = 0; mx < Nx_io; ++mx) {
                const int mxb = mx <= Nx_io / 2 ? mx : mx + (Nx - Nx_io);
                for (int mz = 0; mz < Mz_io; ++mz) u.cmplx(mxb, ny, mz, i) = g.cmplx(mx, ny, mz, i);
            }
    u.makeSpectral_y();
    u.setPadded(true);
    return true;
}

}  // namespace chflow
