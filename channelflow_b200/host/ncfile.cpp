// Reader for the NetCDF-4 field files Channelflow writes (FlowField::writeNetCDF / readNetCDF, reference
// channelflow/flowfield.cpp:3225-3826) without the NetCDF / HDF5 libraries (neither is in this image).  A NetCDF-4 file is
// an HDF5 container; the files of FlowField::writeNetCDF use superblock version 2, version-2 object headers with compact
// link messages, and contiguous little-endian float64 datasets Velocity_X/Y/Z (or Component_i) of shape (Z, Y, X) on the
// I/O grid, plus the scalar global attributes Nx, Ny, Nz, Lx, Lz, a, b of the full grid.  This parses exactly that subset
// and fails loudly on anything else (chunked or compressed layouts, old-style groups).
//
// Writing: FlowField::writeNetCDF produces the same dimensions, grid variables, data variables and global attributes as the
// reference's writer, in the NetCDF *classic* container with 64-bit offsets (CDF-2: a flat big-endian header followed by the
// contiguous variables) instead of the HDF5 one.  nc_open -- which is all FlowField::readNetCDF of stock Channelflow uses --
// detects the container by its magic number, so such a file is read by an unmodified Channelflow; scipy.io.netcdf_file
// reads it too (the independent check in tests/).  The reader below accepts both containers.
#include <cstdint>
#include <cstring>
#include <ctime>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include <unistd.h>

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

// ---------------------------------------------------------------------------------------------- classic container
// CDF-1 / CDF-2 layout (NetCDF classic format specification): magic "CDF" version, numrecs, dim_list, gatt_list, var_list, all
// big-endian and padded to 4 bytes; var entries end with nc_type, vsize, begin (4 or 8 bytes).
struct ClassicVar {
    std::string name;
    std::vector<int> dimids;
    int type = 0;
    uint64_t begin = 0;
};
struct Classic {
    const std::vector<unsigned char>& b;
    size_t p = 4;
    int version;
    std::vector<std::pair<std::string, uint64_t>> dims;
    std::map<std::string, double> gatts;  // numeric scalar attributes
    std::vector<ClassicVar> vars;
    explicit Classic(const std::vector<unsigned char>& bytes) : b(bytes), version(bytes[3]) {}
    uint64_t be(int n) {
        if (p + n > b.size()) cferror("NetCDF reader: truncated classic header");
        uint64_t v = 0;
        for (int i = 0; i < n; ++i) v = (v << 8) | b[p + i];
        p += n;
        return v;
    }
    std::string name() {
        const size_t n = (size_t)be(4);
        if (p + n > b.size()) cferror("NetCDF reader: truncated classic header");
        std::string s((const char*)&b[p], n);
        p += (n + 3) & ~(size_t)3;
        return s;
    }
    static int type_size(int t) { return t == 1 || t == 2 ? 1 : t == 3 ? 2 : t == 4 || t == 5 ? 4 : t == 6 ? 8 : 0; }
    void atts(std::map<std::string, double>* keep) {
        const uint64_t tag = be(4), n = be(4);
        if (tag == 0 && n == 0) return;
        if (tag != 0x0C) cferror("NetCDF reader: attribute list expected in classic header");
        for (uint64_t i = 0; i < n; ++i) {
            const std::string nm = name();
            const int t = (int)be(4);
            const uint64_t ne = be(4);
            const size_t bytes = (size_t)ne * type_size(t);
            if (type_size(t) == 0 || p + bytes > b.size()) cferror("NetCDF reader: bad attribute in classic header");
            if (keep && ne == 1 && (t == 4 || t == 6)) {
                const size_t save = p;
                if (t == 4) (*keep)[nm] = (double)(int32_t)be(4);
                else { const uint64_t u = be(8); double v; memcpy(&v, &u, 8); (*keep)[nm] = v; }
                p = save;
            }
            p += (bytes + 3) & ~(size_t)3;
        }
    }
    void parse() {
        if (version != 1 && version != 2) cferror("NetCDF reader: classic container version " + std::to_string(version) + " is not supported");
        if (be(4) != 0) cferror("NetCDF reader: record variables are not supported");
        uint64_t tag = be(4), n = be(4);
        if (!(tag == 0 && n == 0)) {
            if (tag != 0x0A) cferror("NetCDF reader: dimension list expected in classic header");
            for (uint64_t i = 0; i < n; ++i) { std::string nm = name(); dims.emplace_back(nm, be(4)); }
        }
        atts(&gatts);
        tag = be(4); n = be(4);
        if (!(tag == 0 && n == 0)) {
            if (tag != 0x0B) cferror("NetCDF reader: variable list expected in classic header");
            for (uint64_t i = 0; i < n; ++i) {
                ClassicVar v;
                v.name = name();
                const uint64_t nd = be(4);
                for (uint64_t k = 0; k < nd; ++k) v.dimids.push_back((int)be(4));
                atts(nullptr);
                v.type = (int)be(4);
                be(4);  // vsize
                v.begin = be(version == 1 ? 4 : 8);
                vars.push_back(v);
            }
        }
    }
};

class ClassicWriter {
   public:
    std::vector<unsigned char> h;
    void be(uint64_t v, int n) { for (int i = n - 1; i >= 0; --i) h.push_back((unsigned char)(v >> (8 * i))); }
    void name(const std::string& s) {
        be(s.size(), 4);
        h.insert(h.end(), s.begin(), s.end());
        while (h.size() % 4) h.push_back(0);
    }
    void att_text(const std::string& n, const std::string& v) { name(n); be(2, 4); name(v); }
    void att_int(const std::string& n, int v) { name(n); be(4, 4); be(1, 4); be((uint32_t)v, 4); }
    void att_double(const std::string& n, double v) { uint64_t u; memcpy(&u, &v, 8); name(n); be(6, 4); be(1, 4); be(u, 8); }
};

void put_be_doubles(std::ofstream& os, const double* v, size_t n) {
    std::vector<unsigned char> buf(8 * n);
    for (size_t i = 0; i < n; ++i) {
        uint64_t u;
        memcpy(&u, &v[i], 8);
        for (int k = 0; k < 8; ++k) buf[8 * i + k] = (unsigned char)(u >> (8 * (7 - k)));
    }
    os.write((const char*)buf.data(), (std::streamsize)buf.size());
}

}  // namespace

// FlowField::writeNetCDF (flowfield.cpp:3225-3597): physical values on the I/O grid -- the de-aliased 2(Nx/3) x Ny x 2(Nz/3)
// grid when the field is padded (removePaddedModes, flowfield.cpp:2790-2990), the full grid otherwise -- as variables
// Velocity_X/Y/Z(Z, Y, X) (or the given names / Component_i), grid variables X, Y, Z, and the global attributes of the reference.
void FlowField::writeNetCDF(const std::string& filebase, std::vector<std::string> component_names) const {
    const std::string filename = hasSuffix(filebase, ".nc") ? filebase : filebase + ".nc";
    const int Nx_io = padded() ? 2 * (Nx_ / 3) : Nx_, Nz_io = padded() ? 2 * (Nz_ / 3) : Nz_;
    // physical values on the I/O grid
    FlowField g;
    if (!padded()) {
        g = *this;
        g.makePhysical();
    } else {
        FlowField v(*this);
        v.makeSpectral_xz();
        v.makePhysical_y();
        g = FlowField(Nx_io, Ny_, Nz_io, Nd_, Lx_, Lz_, a_, b_, cfmpi(), Spectral, Physical);
        const int Mz_io = Nz_io / 2 + 1;
        for (int i = 0; i < Nd_; ++i)
            for (int ny = 0; ny < Ny_; ++ny)
                for (int mx = 0; mx < Nx_io; ++mx) {
                    const int mxb = mx <= Nx_io / 2 ? mx : mx + (Nx_ - Nx_io);
                    for (int mz = 0; mz < Mz_io; ++mz) g.cmplx(mx, ny, mz, i) = v.cmplx(mxb, ny, mz, i);
                }
        g.makePhysical_xz();
    }
    std::vector<std::string> var_name;
    if ((int)component_names.size() == Nd_) var_name = component_names;
    else if (Nd_ == 3) var_name = {"Velocity_X", "Velocity_Y", "Velocity_Z"};
    else for (int i = 0; i < Nd_; ++i) var_name.push_back("Component_" + std::to_string(i));

    const uint64_t nvar = (uint64_t)Nx_io * Ny_ * Nz_io * 8;
    if (nvar > 0xFFFFFFFCull) cferror("FlowField::writeNetCDF: a variable of " + std::to_string(nvar) + " bytes does not fit the CDF-2 container");
    if (taskid() != 0) return;

    ClassicWriter w;
    w.h = {'C', 'D', 'F', 2};
    w.be(0, 4);                                   // numrecs
    w.be(0x0A, 4); w.be(3, 4);                    // dimensions, ids 0, 1, 2 = X, Y, Z
    w.name("X"); w.be(Nx_io, 4);
    w.name("Y"); w.be(Ny_, 4);
    w.name("Z"); w.be(Nz_io, 4);
    char tbuf[80] = "";
    { time_t now; time(&now); strftime(tbuf, sizeof tbuf, "%Y-%m-%d %I:%M:%S", localtime(&now)); }
    char host[1024] = "";
    gethostname(host, sizeof host - 1);
    w.be(0x0C, 4); w.be(16, 4);                   // global attributes
    w.att_text("Conventions", "CF-1.0");
    w.att_text("title", "FlowField");
    w.att_text("format_version", "1");
    w.att_text("channelflow_version", "channelflow_b200");
    w.att_text("compiler_version", __VERSION__);
    w.att_text("git_revision", "");
    w.att_text("time", tbuf);
    w.att_text("host_name", host);
    w.att_text("references", "Channelflow is free software: www.channelflow.ch.");
    w.att_int("Nx", Nx_); w.att_int("Ny", Ny_); w.att_int("Nz", Nz_);
    w.att_double("Lx", Lx_); w.att_double("Lz", Lz_); w.att_double("a", a_); w.att_double("b", b_);
    // variables: X(X), Y(Y), Z(Z), then the components with dimensions (Z, Y, X); begin offsets patched below
    w.be(0x0B, 4); w.be(3 + Nd_, 4);
    std::vector<size_t> begin_at;
    std::vector<uint64_t> vsize;
    const char* gname[3] = {"X", "Y", "Z"};
    const uint64_t glen[3] = {(uint64_t)Nx_io, (uint64_t)Ny_, (uint64_t)Nz_io};
    for (int d = 0; d < 3; ++d) {
        w.name(gname[d]); w.be(1, 4); w.be(d, 4); w.be(0, 4); w.be(0, 4); w.be(6, 4); w.be(8 * glen[d], 4);
        begin_at.push_back(w.h.size()); w.be(0, 8); vsize.push_back(8 * glen[d]);
    }
    for (int i = 0; i < Nd_; ++i) {
        w.name(var_name[i]); w.be(3, 4); w.be(2, 4); w.be(1, 4); w.be(0, 4); w.be(0, 4); w.be(0, 4); w.be(6, 4); w.be(nvar, 4);
        begin_at.push_back(w.h.size()); w.be(0, 8); vsize.push_back(nvar);
    }
    uint64_t off = w.h.size();
    for (size_t k = 0; k < begin_at.size(); ++k) {
        for (int i = 0; i < 8; ++i) w.h[begin_at[k] + i] = (unsigned char)(off >> (8 * (7 - i)));
        off += vsize[k];
    }
    std::ofstream os(filename.c_str(), std::ios::out | std::ios::binary);
    if (!os.good()) cferror("FlowField::writeNetCDF: can't open " + filename);
    os.write((const char*)w.h.data(), (std::streamsize)w.h.size());
    std::vector<double> line;
    for (int nx = 0; nx < Nx_io; ++nx) line.push_back(nx * Lx_ / Nx_io);
    put_be_doubles(os, line.data(), line.size());
    const Vector y = ygridpts();
    line.assign(Ny_, 0.0);
    for (int ny = 0; ny < Ny_; ++ny) line[ny] = y[ny];
    put_be_doubles(os, line.data(), line.size());
    line.clear();
    for (int nz = 0; nz < Nz_io; ++nz) line.push_back(nz * Lz_ / Nz_io);
    put_be_doubles(os, line.data(), line.size());
    line.assign(Nx_io, 0.0);
    for (int i = 0; i < Nd_; ++i)
        for (int nz = 0; nz < Nz_io; ++nz)
            for (int ny = 0; ny < Ny_; ++ny) {
                for (int nx = 0; nx < Nx_io; ++nx) line[nx] = g(nx, ny, nz, i);
                put_be_doubles(os, line.data(), line.size());
            }
    if (!os.good()) cferror("FlowField::writeNetCDF: write error on " + filename);
}

// Returns false if the file does not exist; fills u (resized, spectral, padded as the file says) otherwise.
bool read_netcdf_field(const std::string& filename, FlowField& u, CfMPI* cfmpi) {
    std::ifstream is(filename.c_str(), std::ios::in | std::ios::binary);
    if (!is.good()) return false;
    Hdf5 f;
    f.b.assign(std::istreambuf_iterator<char>(is), std::istreambuf_iterator<char>());
    static const unsigned char magic[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    const bool classic = f.b.size() >= 32 && memcmp(f.b.data(), "CDF", 3) == 0;
    std::vector<ObjHeader> comps;
    bool big_endian = false;
    std::map<std::string, double> cl_atts;
    if (classic) {
        Classic c(f.b);
        c.parse();
        big_endian = true;
        cl_atts = c.gatts;
        for (const ClassicVar& v : c.vars) {
            if (v.name == "X" || v.name == "Y" || v.name == "Z") continue;
            ObjHeader o;
            if (v.type != 6 || v.dimids.size() != 3) cferror("NetCDF reader: only float64 variables of shape (Z, Y, X) are supported: " + filename);
            for (int id : v.dimids) {
                if (id < 0 || id >= (int)c.dims.size()) cferror("NetCDF reader: bad dimension id in " + filename);
                o.dims.push_back(c.dims[id].second);
            }
            if (c.dims[v.dimids[0]].first != "Z" || c.dims[v.dimids[1]].first != "Y" || c.dims[v.dimids[2]].first != "X")
                cferror("NetCDF reader: variables must have the dimensions (Z, Y, X): " + filename);
            o.addr = v.begin;
            o.size = 8 * o.dims[0] * o.dims[1] * o.dims[2];
            o.contiguous = true;
            comps.push_back(o);
        }
    }
    auto attr = [&](const char* nm) -> double {
        if (!classic) return find_attr(f, nm);
        if (!cl_atts.count(nm)) cferror(std::string("NetCDF reader: attribute ") + nm + " not found in " + filename);
        return cl_atts.at(nm);
    };
    if (!classic && (f.b.size() < 64 || memcmp(f.b.data(), magic, 8) != 0 || f.b[8] != 2 || f.b[9] != 8))
        cferror("NetCDF reader: " + filename + " is neither a NetCDF classic file nor a NetCDF-4 file with an HDF5 version-2 superblock");
    const ObjHeader root = classic ? ObjHeader() : parse_object(f, f.le(12 + 3 * 8, 8));
    static const char* vel[3] = {"Velocity_X", "Velocity_Y", "Velocity_Z"};
    if (!classic)
        for (const char* nm : vel)
            if (root.links.count(nm)) comps.push_back(parse_object(f, root.links.at(nm)));
    if (comps.empty() && !classic)
        for (int i = 0; root.links.count("Component_" + std::to_string(i)); ++i) comps.push_back(parse_object(f, root.links.at("Component_" + std::to_string(i))));
    if (comps.empty()) cferror("NetCDF reader: no Velocity_X/Y/Z or Component_i variables in " + filename);
    const int Nx = (int)attr("Nx"), Ny = (int)attr("Ny"), Nz = (int)attr("Nz");
    const Real Lx = attr("Lx"), Lz = attr("Lz"), a = attr("a"), b = attr("b");
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
                    const unsigned char* src = base + 8 * ((size_t)nx + (size_t)Nx_io * (ny + (size_t)Ny * nz));
                    double v;
                    if (big_endian) {
                        uint64_t u = 0;
                        for (int k = 0; k < 8; ++k) u = (u << 8) | src[k];
                        memcpy(&v, &u, 8);
                    } else memcpy(&v, src, 8);
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
