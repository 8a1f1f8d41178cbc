// Device-side norms and host-side mean-flow diagnostics (reference diffops.cpp:353-541, 3839-3883).
#include "channelflow/diffops.h"

namespace chflow {

Real L2Norm2(const FlowField& u, bool normalize) {
    Real r = 0;
    cfgpu_check(cfgpu_l2norm2(u.device(), normalize ? 1 : 0, &r), "cfgpu_l2norm2");
    return r;
}
Real L2Norm(const FlowField& u, bool normalize) { return sqrt(L2Norm2(u, normalize)); }
Real L2Norm2_3d(const FlowField& u, bool normalize) {
    double r = 0;
    cfgpu_check(cfgpu_l2norm2_3d(u.device(), normalize ? 1 : 0, &r), "cfgpu_l2norm2_3d");
    return r;
}
Real L2Norm3d(const FlowField& u, bool normalize) { return sqrt(L2Norm2_3d(u, normalize)); }
Real L2Dist2(const FlowField& u, const FlowField& v, bool normalize) {
    Real r = 0;
    cfgpu_check(cfgpu_l2dist2(u.device(), v.device(), normalize ? 1 : 0, &r), "cfgpu_l2dist2");
    return r;
}
Real L2Dist(const FlowField& u, const FlowField& v, bool normalize) { return sqrt(L2Dist2(u, v, normalize)); }
Real L2InnerProduct(const FlowField& u, const FlowField& v, bool normalize) {
    Real r = 0;
    cfgpu_check(cfgpu_l2ip(u.device(), v.device(), normalize ? 1 : 0, &r), "cfgpu_l2ip");
    return r;
}

Real getdPdx(const FlowField& u, Real nu) { return nu * (u.dudy_b() - u.dudy_a()) / (u.b() - u.a()); }
Real getdPdz(const FlowField& u, Real nu) { return nu * (u.dwdy_b() - u.dwdy_a()) / (u.b() - u.a()); }
Real getUbulk(const FlowField& u) {
    Real ubulk = u.profile(0, 0, 0).re.mean();
    if (std::abs(ubulk) < 1e-15) ubulk = 0.0;
    return ubulk;
}
Real getWbulk(const FlowField& u) {
    Real wbulk = u.profile(0, 0, 2).re.mean();
    if (std::abs(wbulk) < 1e-15) wbulk = 0.0;
    return wbulk;
}

}  // namespace chflow
