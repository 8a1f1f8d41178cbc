// chflow::FieldSymmetry / SymmetryList (reference channelflow/symmetry.h:28-96): the discrete-plus-translation symmetries
// of channel flows,  sigma : (u,v,w)(x,y,z) -> s (sx u, sy v, sz w)(sx x + ax Lx, sy y, sz z + az Lz).  Group algebra and
// file formats live on the host; applying sigma to a FlowField is one device kernel (cfgpu_field_symmetry,
// csrc/fieldops.cu:symmetry_kernel).  Not carried: the RealProfile overload and quadraticInterpolate (continuation tools).
#ifndef CHANNELFLOW_SYMMETRY_H
#define CHANNELFLOW_SYMMETRY_H

#include <iostream>
#include <string>

#include "cfbasics/cfarray.h"
#include "cfbasics/cfvector.h"
#include "cfbasics/mathdefs.h"

namespace chflow {

class FlowField;

class FieldSymmetry {
   public:
    FieldSymmetry() = default;
    explicit FieldSymmetry(const std::string& filebase);
    FieldSymmetry(int s) : s_(s) {}
    FieldSymmetry(Real ax, Real az) : ax_(ax), az_(az) {}
    FieldSymmetry(int sx, int sy, int sz, Real ax = 0.0, Real az = 0.0, int s = 1) : s_(s), sx_(sx), sy_(sy), sz_(sz), ax_(ax), az_(az) {}
    FieldSymmetry(bool sx, bool sy, bool sz, Real ax = 0.0, Real az = 0.0, bool s = false)
        : s_(s ? -1 : 1), sx_(sx ? -1 : 1), sy_(sy ? -1 : 1), sz_(sz ? -1 : 1), ax_(ax), az_(az) {}

    FlowField operator()(const FlowField& u) const;
    FieldSymmetry& operator*=(const FieldSymmetry& p);  // (*this) = p * (*this)
    FieldSymmetry& operator*=(Real c) { ax_ *= c; az_ *= c; return *this; }

    int s() const { return s_; }
    int sx() const { return sx_; }
    int sy() const { return sy_; }
    int sz() const { return sz_; }
    Real ax() const { return ax_; }
    Real az() const { return az_; }
    int s(int i) const { return i == 0 ? sx_ : i == 1 ? sy_ : sz_; }
    int sign(int i) const { return s(i); }
    void save(const std::string& filebase, std::ios::openmode openflag = std::ios::out) const;
    bool isIdentity() const { return s_ == 1 && sx_ == 1 && sy_ == 1 && sz_ == 1 && ax_ == 0.0 && az_ == 0.0; }

   private:
    int s_ = 1, sx_ = 1, sy_ = 1, sz_ = 1;
    Real ax_ = 0.0, az_ = 0.0;
};

bool operator==(const FieldSymmetry& p, const FieldSymmetry& q);
bool operator!=(const FieldSymmetry& p, const FieldSymmetry& q);
std::istream& operator>>(std::istream& is, FieldSymmetry& s);
std::ostream& operator<<(std::ostream& os, const FieldSymmetry& s);
FieldSymmetry operator*(const FieldSymmetry& p, const FieldSymmetry& q);  // (p*q)(u) = p(q(u))
FieldSymmetry inverse(const FieldSymmetry& s);
FlowField operator*(const FieldSymmetry& s, const FlowField& u);

class SymmetryList : public cfarray<FieldSymmetry> {
   public:
    SymmetryList() {}
    explicit SymmetryList(int n) : cfarray<FieldSymmetry>(n) {}
    explicit SymmetryList(const std::string& filebase);
    void save(const std::string& filebase) const;
};
std::ostream& operator<<(std::ostream& os, const SymmetryList& s);

void project(const FieldSymmetry& s, const FlowField& u, FlowField& Pu);  // Pu = (1 + s)/2 u
FlowField project(const FieldSymmetry& s, FlowField& u);
void project(const cfarray<FieldSymmetry>& s, const FlowField& u, FlowField& Pu);  // Pu = prod_n (1 + s[n])/2 u
FlowField project(const cfarray<FieldSymmetry>& s, FlowField& u);

}  // namespace chflow
#endif
