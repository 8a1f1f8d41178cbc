// FieldSymmetry algebra, file formats and projections (reference channelflow/symmetry.cpp); sigma * FlowField on the device.
#include "channelflow/symmetry.h"

#include <fstream>
#include <iomanip>

#include "cfgpu.h"
#include "channelflow/flowfield.h"

using namespace std;

namespace chflow {

FieldSymmetry::FieldSymmetry(const string& filebase) {
    ifstream is;
    const string filename = ifstreamOpen(is, filebase, ".asc");
    if (!is) cferror("FieldSymmetry(filebase) : can't open file " + filebase + " or " + filebase + ".asc");
    // an optional "% ..." header line, then  s sx sy sz ax az
    string line;
    while (is.peek() == '%' || is.peek() == '\n') getline(is, line);
    is >> *this;
}

void FieldSymmetry::save(const string& filebase, ios::openmode openflag) const {
    ofstream os(appendSuffix(filebase, ".asc").c_str(), openflag);
    os << setprecision(17) << *this << endl;
}

// composition law: (p q)(u) = p(q(u)): signs multiply, q's shift is reflected by p before p's shift is added
FieldSymmetry& FieldSymmetry::operator*=(const FieldSymmetry& p) { return *this = p * (*this); }

FieldSymmetry operator*(const FieldSymmetry& p, const FieldSymmetry& q) {
    return FieldSymmetry(p.sx() * q.sx(), p.sy() * q.sy(), p.sz() * q.sz(), p.ax() + p.sx() * q.ax(), p.az() + p.sz() * q.az(), p.s() * q.s());
}
FieldSymmetry inverse(const FieldSymmetry& s) { return FieldSymmetry(s.sx(), s.sy(), s.sz(), -s.ax(), -s.az(), s.s()); }
bool operator==(const FieldSymmetry& p, const FieldSymmetry& q) {
    return p.s() == q.s() && p.sx() == q.sx() && p.sy() == q.sy() && p.sz() == q.sz() && p.ax() == q.ax() && p.az() == q.az();
}
bool operator!=(const FieldSymmetry& p, const FieldSymmetry& q) { return !(p == q); }

ostream& operator<<(ostream& os, const FieldSymmetry& s) {
    for (int v : {s.s(), s.sx(), s.sy(), s.sz()}) os << setw(3) << right << v;
    return os << resetiosflags(ios::adjustfield) << '\t' << s.ax() << '\t' << s.az();
}
istream& operator>>(istream& is, FieldSymmetry& sigma) {
    int s = 1, sx = 1, sy = 1, sz = 1;
    Real ax = 0, az = 0;
    is >> s >> sx >> sy >> sz >> ax >> az;
    sigma = FieldSymmetry(sx, sy, sz, ax, az, s);
    return is;
}

FlowField FieldSymmetry::operator()(const FlowField& u) const {
    FlowField v(u);
    v *= *this;
    return v;
}
FlowField operator*(const FieldSymmetry& s, const FlowField& u) { return s(u); }

FlowField& FlowField::operator*=(const FieldSymmetry& sigma) {
    if (sigma.isIdentity()) return *this;
    const fieldstate xzs = xzstate_, ys = ystate_;
    makeState(Spectral, Spectral);
    cfgpu_check(cfgpu_field_symmetry(device_mut(), sigma.s(), sigma.sx(), sigma.sy(), sigma.sz(), sigma.ax(), sigma.az()), "FlowField *= FieldSymmetry");
    makeState(xzs, ys);
    return *this;
}

FlowField& FlowField::project(const FieldSymmetry& sigma) {  // flowfield.cpp:1095-1266: P u = (u + sigma u)/2
    if (sigma.isIdentity()) return *this;
    FlowField su(*this);
    su *= sigma;
    add(1.0, su);
    *this *= 0.5;
    return *this;
}
FlowField& FlowField::project(const cfarray<FieldSymmetry>& sigma) {
    for (int n = 0; n < sigma.length(); ++n) project(sigma[n]);
    return *this;
}

// ---- lists: "% N" header, one symmetry per line
SymmetryList::SymmetryList(const string& filebase) {
    ifstream is;
    const string filename = ifstreamOpen(is, filebase, ".asc");
    if (!is) cferror("SymmetryList(filebase) : can't open file " + filebase + " or " + filebase + ".asc");
    char c = 0;
    int N = 0;
    is >> c;
    if (c != '%') cferror("SymmetryList(filebase) : bad header in file " + filename);
    is >> N;
    resize(N);
    for (int n = 0; n < N; ++n) is >> (*this)[n];
}
void SymmetryList::save(const string& filebase) const {
    ofstream os(appendSuffix(filebase, ".asc").c_str());
    os << setprecision(17) << "% " << length() << '\n' << *this;
}
ostream& operator<<(ostream& os, const SymmetryList& s) {
    for (int n = 0; n < s.length(); ++n) os << s[n] << '\n';
    return os;
}

// ---- projections onto the sigma-symmetric subspace
void project(const FieldSymmetry& s, const FlowField& u, FlowField& Pu) {
    FlowField su = s(u);
    su += u;
    su *= 0.5;
    Pu = su;
}
FlowField project(const FieldSymmetry& s, FlowField& u) {
    FlowField Pu;
    project(s, u, Pu);
    return Pu;
}
void project(const cfarray<FieldSymmetry>& s, const FlowField& u, FlowField& Pu) {
    FlowField acc(u);
    for (int n = 0; n < s.length(); ++n) {
        FlowField next;
        project(s[n], acc, next);
        acc = next;
    }
    Pu = acc;
}
FlowField project(const cfarray<FieldSymmetry>& s, FlowField& u) {
    FlowField Pu;
    project(s, u, Pu);
    return Pu;
}

}  // namespace chflow
