// Norms, differential operators and pointwise products of FlowFields: the free functions of the reference's
// channelflow/diffops.h.  Every function runs on the device (include/cfgpu.h: cfgpu_l2*, cfgpu_field_diffop,
// cfgpu_field_pointwise, cfgpu_bcnorm2); the basis-set projections (BasisFunc / RealProfile) of the reference header
// are outside this package's scope.
#ifndef CFB200_DIFFOPS_H
#define CFB200_DIFFOPS_H
#include <string>
#include <vector>

#include "cfbasics/cfbasics.h"
#include "cfbasics/cfvector.h"
#include "cfbasics/mathdefs.h"
#include "channelflow/chebyshev.h"
#include "channelflow/flowfield.h"

namespace chflow {

Real L2Norm(const FlowField& f, bool normalize = true);
Real L2Norm2(const FlowField& f, bool normalize = true);
Real L2Dist(const FlowField& f, const FlowField& g, bool normalize = true);
Real L2Dist2(const FlowField& f, const FlowField& g, bool normalize = true);
Real chebyNorm(const FlowField& f, bool normalize = true);
Real chebyNorm2(const FlowField& f, bool normalize = true);
Real chebyDist(const FlowField& f, const FlowField& g, bool normalize = true);
Real chebyDist2(const FlowField& f, const FlowField& g, bool normalize = true);
Real bcNorm(const FlowField& f, bool normalize = true);
Real bcNorm2(const FlowField& f, bool normalize = true);
Real bcDist(const FlowField& f, const FlowField& g, bool normalize = true);
Real bcDist2(const FlowField& f, const FlowField& g, bool normalize = true);
Real divNorm(const FlowField& f, bool normalize = true);
Real divNorm2(const FlowField& f, bool normalize = true);
Real divDist(const FlowField& f, const FlowField& g, bool normalize = true);
Real divDist2(const FlowField& f, const FlowField& g, bool normalize = true);
Real L2Norm(const FlowField& f, int kxmax, int kzmax, bool normalize = true);
Real L2Norm2(const FlowField& f, int kxmax, int kzmax, bool normalize = true);
Real L2Dist(const FlowField& f, const FlowField& g, int kxmax, int kzmax, bool normalize = true);
Real L2Dist2(const FlowField& f, const FlowField& g, int kxmax, int kzmax, bool normalize = true);
Real L2InnerProduct(const FlowField& f, const FlowField& g, int kxmax, int kzmax, bool normalize = true);
Real L2InnerProduct(const FlowField& f, const FlowField& g, bool normalize = true);
inline Real L2IP(const FlowField& f, const FlowField& g, bool normalize = true) { return L2InnerProduct(f, g, normalize); }
inline Real L2IP(const FlowField& f, const FlowField& g, int kxmax, int kzmax, bool normalize = true) {
    return L2InnerProduct(f, g, kxmax, kzmax, normalize);
}
Real dissipation(const FlowField& f, bool normalize = true);     // 1/(LxLyLz) int |grad u|^2
Real wallshear(const FlowField& f, bool normalize = true);       // mean of the two walls
Real wallshearLower(const FlowField& f, bool normalize = true);
Real wallshearUpper(const FlowField& f, bool normalize = true);
Real L2Norm2_3d(const FlowField& f, bool normalize = true);      // without the kx = 0 modes
Real L2Norm3d(const FlowField& f, bool normalize = true);
Real L2Norm_uvw(const FlowField& u, const bool ux, const bool uy, const bool uz);
Real Ecf(const FlowField& u);

FlowField xdiff(const FlowField& f, int n = 1);
FlowField ydiff(const FlowField& f, int n = 1);
FlowField zdiff(const FlowField& f, int n = 1);
FlowField diff(const FlowField& f, int i, int n);
FlowField diff(const FlowField& f, int nx, int ny, int nz);
FlowField grad(const FlowField& f);
FlowField lapl(const FlowField& f);
FlowField curl(const FlowField& f);
FlowField norm(const FlowField& f);
FlowField norm2(const FlowField& f);
FlowField div(const FlowField& f);
FlowField cross(const FlowField& f, const FlowField& g);
FlowField outer(const FlowField& f, const FlowField& g);
FlowField dot(const FlowField& f, const FlowField& g);
FlowField energy(const FlowField& u);
FlowField energy(const FlowField& u, ChebyCoeff& U);

void xdiff(const FlowField& f, FlowField& dfdx, int n = 1);
void ydiff(const FlowField& f, FlowField& dfdy, int n = 1);
void zdiff(const FlowField& f, FlowField& dfdz, int n = 1);
void diff(const FlowField& f, FlowField& df, int i, int n);
void diff(const FlowField& f, FlowField& df, int nx, int ny, int nz);
void grad(const FlowField& f, FlowField& grad_f);  // 1 -> 3 or 3 -> 9 components, grad_f[3i+j] = d f_i/d x_j
void lapl(const FlowField& f, FlowField& lapl_f);
void curl(const FlowField& f, FlowField& curl_f);
void norm(const FlowField& f, FlowField& norm_f);
void norm2(const FlowField& f, FlowField& norm2_f);
void div(const FlowField& f, FlowField& divf, const fieldstate finalstate = Spectral);
void cross(const FlowField& f, const FlowField& g, FlowField& f_cross_g, const fieldstate finalstate = Spectral);
void outer(const FlowField& f, const FlowField& g, FlowField& fg);
void dot(const FlowField& f, const FlowField& g, FlowField& f_dot_g);
void energy(const FlowField& u, FlowField& e);
void energy(const FlowField& u, const ChebyCoeff& U, FlowField& e);
void dotgrad(const FlowField& u, const FlowField& v, FlowField& u_dotgrad_v, FlowField& tmp);
FlowField dotgrad(const FlowField& u, const FlowField& v, FlowField& tmp);

// random divergence-free, no-slip profiles of one Fourier mode on the libc drand48 stream (diffops.cpp:969-1243): the
// construction rule of FlowField::addPerturbations and tools/randomfield.cpp
void randomUprofile(ComplexChebyCoeff& u, Real mag, Real spectralDecay);
void randomVprofile(ComplexChebyCoeff& v, Real mag, Real spectralDecay);
void randomProfile(ComplexChebyCoeff& u, ComplexChebyCoeff& v, ComplexChebyCoeff& w, int kx, int kz, Real Lx, Real Lz, Real mag,
                   Real spectralDecay);
void chebyUprofile(ComplexChebyCoeff& u, int n, Real decay);
void chebyVprofile(ComplexChebyCoeff& v, int n, Real decay);

Real getdPdx(const FlowField& u, Real nu);
Real getdPdz(const FlowField& u, Real nu);
Real getUbulk(const FlowField& u);
Real getWbulk(const FlowField& u);

std::string fieldstats_t(const FlowField& u, Real t);
std::string fieldstatsheader_t(const std::string tname = "t");
std::string fieldstats(const FlowField& u);
std::string fieldstatsheader();

// The nonlinear term of a (total) velocity field in its different forms (diffops.cpp:2852-3365): one device pipeline each
// (cfgpu_nse_nonlinear with a zero base flow); tmp is scratch of the reference's host algorithm and is not used.
void rotationalNL(const FlowField& u, FlowField& f, FlowField& tmp, const fieldstate finalstate = Spectral);
void convectionNL(const FlowField& u, FlowField& f, FlowField& tmp, const fieldstate finalstate = Spectral);
void divergenceNL(const FlowField& u, FlowField& f, FlowField& tmp, const fieldstate finalstate = Spectral);
void skewsymmetricNL(const FlowField& u, FlowField& f, FlowField& tmp, const fieldstate finalstate = Spectral);
void linearizedNL(const FlowField& u, const ChebyCoeff& U, const ChebyCoeff& W, FlowField& f, const fieldstate finalstate = Spectral);

}  // namespace chflow
#endif
