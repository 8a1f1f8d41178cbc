// Utilities shared by the Channelflow programs (the part of the reference's channelflow/utilfuncs.h that the DNS-side
// programs use): run records, boundary-condition fixes of Chebyshev profiles, a field time series for interpolation,
// base-flow construction from command-line flags.
#ifndef CFB200_UTILFUNCS_H
#define CFB200_UTILFUNCS_H
#include <unistd.h>

#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>

#include "cfbasics/arglist.h"
#include "cfbasics/cfarray.h"
#include "cfbasics/mathdefs.h"
#include "channelflow/config.h"
#include "channelflow/diffops.h"
#include "channelflow/nse.h"

namespace chflow {

void WriteProcessInfo(int argc, char* argv[], std::string filename = "processinfo", std::ios::openmode mode = std::ios::out);

// (fixDiri / fixDiriMean are declared with FlowField's vector maps in channelflow/flowfield.h)
void fixDiriNeum(ChebyCoeff& f);
void fixDiriNeum(ComplexChebyCoeff& f);

// the last N fields of a time series, interpolated in time with an (N-1)th-order polynomial
class FieldSeries {
   public:
    FieldSeries();
    FieldSeries(int N);
    void push(const FlowField& f, Real t);
    void interpolate(FlowField& f, Real t) const;
    bool full() const;

   private:
    cfarray<Real> t_;
    cfarray<FlowField> f_;
    int emptiness_;
};

Real tFromFilename(const std::string filename);                   // "u12.5.ff" -> 12.5
bool comparetimes(const std::string& s0, const std::string& s1);  // ordering of such file names
void channelflowVersion(int& major, int& minor, int& update);

DNSFlags setBaseFlowFlags(ArgList& args, std::string& Uname, std::string& Wname);
std::vector<ChebyCoeff> baseFlow(int Ny, Real a, Real b, DNSFlags& flags, std::string Uname, std::string Wname);

}  // namespace chflow
#endif
