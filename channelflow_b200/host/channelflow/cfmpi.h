// Process grid of a run: one process per B200.  Same public surface as the reference's CfMPI (channelflow/cfmpi.h:34-118,
// cfmpi.cpp:66-127), but the "communicator" underneath is the device layer's (cfgpu_comm_*, NCCL over NVLink): the
// decomposition is always the kx-slab / y-slab pair of include/cfgpu.h, so the np0 x np1 pencil grid a caller asks for is
// recorded and reported, not used to lay out data.  cfMPI_Init launches nothing: the processes are started by an
// external launcher (torchrun / mpirun / srun: RANK or OMPI_COMM_WORLD_RANK or PMI_RANK, WORLD_SIZE..., MASTER_ADDR,
// MASTER_PORT) and rendezvous here to share the NCCL id.
#ifndef CFB200_CFMPI_H
#define CFB200_CFMPI_H
#include <cstdlib>

#include "cfbasics/mathdefs.h"

namespace chflow {

int cfMPI_Init(int* argc, char*** argv);  // joins the device communicator when launched with a world size > 1
int cfMPI_Finalize();

class CfMPI {
   protected:
    CfMPI(int nproc0, int nproc1);
    CfMPI(const CfMPI&);
    CfMPI& operator=(const CfMPI&);

   public:
    static CfMPI& getInstance(int nproc0 = 0, int nproc1 = 0) {
        static CfMPI instance(nproc0, nproc1);
        return instance;
    }
    MPI_Comm comm0 = 0, comm1 = 0, comm_world = 0;  // placeholders: collectives run inside the device layer

    int nproc0() const { return nproc0_; }
    int nproc1() const { return nproc1_; }
    int taskid() const { return taskid_; }
    int taskid_world() const { return taskid_; }
    int numtasks() const { return numtasks_; }
    int color0() const { return taskid_ / nproc0_; }
    int key0() const { return taskid_ % nproc0_; }
    int color1() const { return taskid_ % nproc0_; }
    int key1() const { return taskid_ / nproc0_; }
    int usempi_ = 0;

   private:
    int nproc0_, nproc1_, taskid_, numtasks_;
};

class CfMPI_single : public CfMPI {
    using CfMPI::CfMPI;

   public:
    static CfMPI_single& getInstance() {
        static CfMPI_single instance(1, 1);
        return instance;
    }
};

// rank / world size this process was launched with (environment of the launcher)
int launch_rank();
int launch_world_size();

}  // namespace chflow
#endif
