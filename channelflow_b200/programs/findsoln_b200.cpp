// findsoln_b200 -- Newton-Krylov-hookstep search for (relative) equilibria of plane shear flows with every state vector on
// the GPU: the fixed-T subset of the reference's programs/findsoln.cpp (`findsoln -eqb [-xrel] [-zrel] -T <T> [-sigma file]`)
// on channelflow/devicesearch.h.  Same option names and defaults as the reference (programs/findsoln.cpp:20-58,
// nsolver/newtonalgorithm.cpp: NewtonSearchFlags(ArgList&)); unknown-period (-orb) and multishooting searches are rejected.
// Writes <outdir>/ubest.ff, <outdir>/sigmabest.asc and <outdir>/convergence.asc.
#include <fstream>
#include <iomanip>
#include <iostream>

#include "cfbasics/arglist.h"
#include "channelflow/devicesearch.h"
#include "channelflow/diffops.h"
#include "channelflow/dnsflags.h"
#include "channelflow/flowfield.h"
#include "channelflow/symmetry.h"

using namespace std;
using namespace chflow;

int main(int argc, char* argv[]) {
    cfMPI_Init(&argc, &argv);
    {
        ArgList args(argc, argv, "find an invariant solution of plane Couette or channel flow with Newton-Krylov-hookstep search (device resident)");
        args.section("Newton-Krylov-hookstep search");
        DeviceSearchFlags sf;
        const bool eqb = args.getflag("-eqb", "--equilibrium", "search for equilibrium or relative equilibrium (trav wave)");
        const bool orb = args.getflag("-orb", "--periodicorbit", "search for periodic orbit (not available in the device search)");
        const bool xrel = args.getflag("-xrel", "--xrelative", "search over x phase shift for relative solution");
        const bool zrel = args.getflag("-zrel", "--zrelative", "search over z phase shift for relative solution");
        sf.epsSearch = args.getreal("-es", "--epsSearch", 1e-13, "stop search if L2Norm(s f^T(u) - u) < epsEQB");
        sf.epsKrylov = args.getreal("-ek", "--epsKrylov", 1e-14, "min. condition # of Krylov vectors");
        sf.epsDx = args.getreal("-edx", "--epsDxLinear", 1e-7, "relative size of dx to x in linearization");
        sf.epsGMRES = args.getreal("-eg", "--epsGMRES", 1e-3, "stop GMRES iteration when Ax=b residual is < this");
        sf.epsGMRESf = args.getreal("-egf", "--epsGMRESfinal", 0.05, "accept final GMRES iterate if residual is < this");
        sf.centdiff = args.getflag("-cd", "--centerdiff", "centered differencing to estimate differentials");
        sf.Nnewton = args.getint("-Nn", "--Nnewton", 20, "max number of Newton steps");
        sf.Ngmres = args.getint("-Ng", "--Ngmres", 120, "max number of GMRES iterations per restart");
        sf.Nhook = args.getint("-Nh", "--Nhook", 20, "max number of hookstep iterations per Newton");
        sf.delta = args.getreal("-d", "--delta", 0.01, "initial radius of trust region");
        sf.deltaMin = args.getreal("-dmin", "--deltaMin", 1e-12, "stop if radius of trust region gets this small");
        sf.deltaMax = args.getreal("-dmax", "--deltaMax", 0.1, "maximum radius of trust region");
        sf.lambdaMin = args.getreal("-lmin", "--lambdaMin", 0.2, "minimum delta shrink rate");
        sf.lambdaMax = args.getreal("-lmax", "--lambdaMax", 1.5, "maximum delta expansion rate");
        sf.improvReq = args.getreal("-irq", "--improveReq", 1e-3, "reduce delta and recompute hookstep if improvement is worse than this fraction of what we'd expect from gradient");
        sf.improvOk = args.getreal("-iok", "--improveOk", 0.10, "accept step and keep same delta if improvement is better than this fraction of quadratic model");
        sf.improvGood = args.getreal("-igd", "--improveGood", 0.75, "accept step and increase delta if improvement is better than this fraction of quadratic model");
        const string outdir = pathfix(args.getpath("-o", "--outdir", "./", "output directory"));

        DNSFlags dnsflags(args);
        TimeStep dt(dnsflags);

        args.section("Program options");
        const string sigmastr = args.getstr("-sigma", "--sigma", "", "file containing sigma of sigma f^T(u) - u = 0 (default == identity)");
        const string uname = args.getstr(1, "<flowfield>", "initial guess for the solution");
        args.check();
        args.save(outdir);
        if (orb || !eqb) cferror("findsoln_b200: only -eqb searches (fixed integration time T) run on the device; use stock findsoln for -orb");

        FlowField u(uname);
        FieldSymmetry sigma;
        if (!sigmastr.empty()) sigma = FieldSymmetry(sigmastr);
        dnsflags.verbosity = Silent;
        cout << setprecision(16) << " 1/nu == " << 1 / dnsflags.nu << "\nsigma == " << sigma << "\n    T == " << dnsflags.T << "\n   dt == " << dt
             << "\nDNSFlags == " << dnsflags << '\n' << endl;

        u.makeSpectral();
        DeviceDSI dsi(u, dnsflags, dt, sigma, dnsflags.T, /*Tnormalize*/ true, xrel, zrel);
        DeviceVector x;
        dsi.makeVector(u, x);
        const DeviceSearchResult r = hookstepSearch(dsi, x, sf);
        dsi.extractVector(x, u);
        u.save(outdir + "ubest");
        dsi.sigma().save(outdir + "sigmabest");
        {
            ofstream os((outdir + "convergence.asc").c_str());
            os << setprecision(16) << "% L2Norm(G) per Newton-hookstep step\n";
            for (Real g : r.history) os << g << '\n';
        }
        cout << setprecision(8) << (r.converged ? "converged" : "stopped") << ": L2Norm(G) == " << r.residual << " after " << r.newtonSteps
             << " Newton steps, " << r.fevals << " DNS integrations, " << r.gmresIterations << " GMRES iterations\nsigma == " << setprecision(17)
             << dsi.sigma() << endl;
    }
    cfMPI_Finalize();
    return 0;
}
