// Norms and mean-flow diagnostics of FlowFields (the hot-path subset of the reference's channelflow/diffops.h).
#ifndef CFB200_DIFFOPS_H
#define CFB200_DIFFOPS_H
#include "channelflow/flowfield.h"

namespace chflow {
Real L2Norm2(const FlowField& u, bool normalize = true);
Real L2Norm2_3d(const FlowField& u, bool normalize = true);  // without the kx = 0 modes (diffops.cpp:700-740)
Real L2Norm3d(const FlowField& u, bool normalize = true);
Real L2Norm(const FlowField& u, bool normalize = true);
Real L2Dist2(const FlowField& u, const FlowField& v, bool normalize = true);
Real L2Dist(const FlowField& u, const FlowField& v, bool normalize = true);
Real L2InnerProduct(const FlowField& u, const FlowField& v, bool normalize = true);
Real getdPdx(const FlowField& u, Real nu);
Real getdPdz(const FlowField& u, Real nu);
Real getUbulk(const FlowField& u);
Real getWbulk(const FlowField& u);
}  // namespace chflow
#endif
