// Process-grid singleton.  The reference's CfMPI (channelflow/cfmpi.h:38-97) wraps MPI communicators; here one
// process drives one B200 and multi-GPU runs are one process per GPU (torch.distributed / NCCL launches them), so
// this class only reports the grid and this process' place in it.
#ifndef CFB200_CFMPI_H
#define CFB200_CFMPI_H
#include <cstdlib>

namespace chflow {

class CfMPI {
   public:
    static CfMPI& getInstance(int nproc0 = 0, int nproc1 = 0) {
        static CfMPI inst(nproc0, nproc1);
        return inst;
    }
    int nproc0() const { return nproc0_; }
    int nproc1() const { return nproc1_; }
    int taskid() const { return taskid_; }
    int taskid_world() const { return taskid_; }
    int numtasks() const { return numtasks_; }
    int key0() const { return taskid_ % nproc0_; }
    int color0() const { return taskid_ / nproc0_; }

   private:
    CfMPI(int np0, int np1) {
        const char* r = std::getenv("RANK");
        const char* w = std::getenv("WORLD_SIZE");
        taskid_ = r ? std::atoi(r) : 0;
        numtasks_ = w ? std::atoi(w) : 1;
        nproc0_ = np0 > 0 ? np0 : 1;
        nproc1_ = np1 > 0 ? np1 : numtasks_ / nproc0_;
        if (nproc1_ < 1) nproc1_ = 1;
    }
    int nproc0_, nproc1_, taskid_, numtasks_;
};

inline void cfMPI_Init(int*, char***) {}
inline void cfMPI_Finalize() {}
inline int mpirank() { return CfMPI::getInstance().taskid(); }

}  // namespace chflow
#endif
