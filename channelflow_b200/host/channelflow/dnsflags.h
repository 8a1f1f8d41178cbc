// DNSFlags / TimeStep and the DNS enums -- same names, fields, defaults and semantics as the reference's
// channelflow/dnsflags.h:24-198 (host-side control; nothing here touches the GPU).
#ifndef CFB200_DNSFLAGS_H
#define CFB200_DNSFLAGS_H
#include <iostream>
#include <string>
#include <vector>

#include "cfbasics/arglist.h"
#include "cfbasics/cfvector.h"
#include "cfbasics/mathdefs.h"
#include "channelflow/chebyshev.h"
#include "channelflow/flowfield.h"
#include "channelflow/symmetry.h"

namespace chflow {

enum VelocityScale { WallScale, ParabolicScale };
enum BaseFlow { ZeroBase, LinearBase, ParabolicBase, LaminarBase, SuctionBase, ArbitraryBase };
enum MeanConstraint { PressureGradient, BulkVelocity };
enum TimeStepMethod { CNFE1, CNAB2, CNRK2, SMRK2, SBDF1, SBDF2, SBDF3, SBDF4 };
enum NonlinearMethod { Rotational, Convection, Divergence, SkewSymmetric, Alternating, Alternating_, LinearAboutProfile };
enum Dealiasing { NoDealiasing, DealiasXZ, DealiasY, DealiasXYZ };
enum Verbosity { Silent, PrintTime, PrintTicks, VerifyTauSolve, PrintAll };

VelocityScale s2velocityscale(const std::string& s);
BaseFlow s2baseflow(const std::string& s);
MeanConstraint s2constraint(const std::string& s);
TimeStepMethod s2stepmethod(const std::string& s);
NonlinearMethod s2nonlmethod(const std::string& s);
Dealiasing s2dealiasing(const std::string& s);
Verbosity s2verbosity(const std::string& s);
std::string baseflow2string(BaseFlow bf);
std::string constraint2string(MeanConstraint mc);
std::string stepmethod2string(TimeStepMethod ts);
std::string nonlmethod2string(NonlinearMethod nm);
std::string dealiasing2string(Dealiasing d);
std::ostream& operator<<(std::ostream& os, BaseFlow b);
std::ostream& operator<<(std::ostream& os, MeanConstraint m);
std::ostream& operator<<(std::ostream& os, TimeStepMethod t);
std::ostream& operator<<(std::ostream& os, NonlinearMethod n);
std::ostream& operator<<(std::ostream& os, Dealiasing d);

std::ostream& operator<<(std::ostream& os, VelocityScale v);
std::ostream& operator<<(std::ostream& os, Verbosity v);

// A body force f(x,y,z,t).  As in the reference (dnsflags.h:67-78) the time steppers never evaluate it: the class exists
// so that programs which set DNSFlags::bodyforce compile and run.
class BodyForce {
   public:
    BodyForce() {}
    virtual ~BodyForce() = default;
    Vector operator()(Real x, Real y, Real z, Real t);
    void eval(Real t, FlowField& f);
    virtual void eval(Real x, Real y, Real z, Real t, Real& fx, Real& fy, Real& fz);
    virtual bool isOn(Real t);
};

class DNSFlags {
   public:
    DNSFlags(Real nu = 0.0025, Real dPdx = 0.0, Real dPdz = 0.0, Real Ubulk = 0.0, Real Wbulk = 0.0, Real Uwall = 1.0,
             Real ulowerwall = 0.0, Real uupperwall = 0.0, Real wlowerwall = 0.0, Real wupperwall = 0.0,
             Real theta = 0.0, Real Vsuck = 0.0, Real rotation = 0.0, Real t0 = 0.0, Real T = 20.0, Real dT = 1.0,
             Real dt = 0.03125, bool variabledt = true, Real dtmin = 0.001, Real dtmax = 0.2, Real CFLmin = 0.4,
             Real CFLmax = 0.6, Real symmetryprojectioninterval = 100.0, BaseFlow baseflow = LaminarBase,
             MeanConstraint constraint = PressureGradient, TimeStepMethod timestepping = SBDF3,
             TimeStepMethod initstepping = SMRK2, NonlinearMethod nonlinearity = Rotational,
             Dealiasing dealiasing = DealiasXZ, BodyForce* bodyforce = 0, bool taucorrection = true,
             Verbosity verbosity = PrintTicks, std::ostream* logstream = &std::cout);
    DNSFlags(ArgList& args, const bool laurette = false);  // the command-line options of the reference's programs
    virtual ~DNSFlags() = default;
    void args2BC(ArgList& args);
    void args2numerics(ArgList& args, const bool laurette = false);
    virtual void save(const std::string& outdir = "") const;  // dnsflags.txt
    virtual void load(int taskid, const std::string indir);

    bool dealias_xz() const { return dealiasing == DealiasXZ || dealiasing == DealiasXYZ; }
    bool dealias_y() const { return dealiasing == DealiasY || dealiasing == DealiasXYZ; }

    BaseFlow baseflow;
    MeanConstraint constraint;
    TimeStepMethod timestepping;
    TimeStepMethod initstepping;
    NonlinearMethod nonlinearity;
    Dealiasing dealiasing;
    BodyForce* bodyforce;
    bool taucorrection;

    Real nu, Vsuck, rotation, theta, dPdx, dPdz, Ubulk, Wbulk, Uwall;
    Real ulowerwall, uupperwall, wlowerwall, wupperwall;
    Real t0, T, dT, dt;
    bool variabledt;
    Real dtmin, dtmax, CFLmin, CFLmax;
    int symmetryprojectioninterval;
    Verbosity verbosity;
    std::ostream* logstream;
    cfarray<FieldSymmetry> symmetries;  // restrict u(t) to these symmetries (projection every symmetryprojectioninterval)
    std::string symmetries_file;  // -symms <file>: generators of the isotropy group to project onto (see symmetry.h)
};

std::ostream& operator<<(std::ostream& os, const DNSFlags& flags);

// Keeps dt an integer fraction of dT and CFL inside [CFLmin, CFLmax] (reference dnsflags.cpp:749-975).
class TimeStep {
   public:
    TimeStep();
    TimeStep(Real dt, Real dtmin, Real dtmax, Real dT, Real CFLmin, Real CFLmax, bool variable = true);
    TimeStep(DNSFlags& flags);

    bool adjust(Real CFL, bool verbose = true, std::ostream& os = std::cout);
    bool adjustToMiddle(Real CFL, bool verbose = true, std::ostream& os = std::cout);
    bool adjust(Real a, Real a_max, bool verbose = true, std::ostream& os = std::cout);
    bool adjustToDesired(Real a, Real a_des, bool verbose = true, std::ostream& os = std::cout);
    bool adjust_for_T(Real T, bool verbose = true, std::ostream& os = std::cout);

    int n() const { return n_; }
    int N() const { return N_; }
    Real dt() const { return dt_; }
    Real dtmin() const { return dtmin_; }
    Real dtmax() const { return dtmax_; }
    Real dT() const { return dT_; }
    Real T() const { return T_; }
    Real CFL() const { return CFL_; }
    Real CFLmin() const { return CFLmin_; }
    Real CFLmax() const { return CFLmax_; }
    bool variable() const { return variable_; }
    operator Real() const { return dT_ / n_; }

   private:
    int n_, N_;
    Real dt_, dtmin_, dtmax_, dT_, T_, CFLmin_, CFL_, CFLmax_;
    bool variable_;
};

std::ostream& operator<<(std::ostream& os, const TimeStep& ts);

}  // namespace chflow
#endif
