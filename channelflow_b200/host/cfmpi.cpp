// CfMPI over the device communicator (see channelflow/cfmpi.h).  The only host-side communication is the rendezvous that
// hands the 128-byte NCCL id from rank 0 to the other ranks at start-up (a TCP exchange on MASTER_ADDR:MASTER_PORT+1, the
// convention of the usual launchers); everything afterwards travels GPU to GPU.
#include "channelflow/cfmpi.h"

#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <unistd.h>

#include <cstring>
#include <string>

#include "channelflow/flowfield.h"

namespace chflow {

static int env_int(std::initializer_list<const char*> names, int dflt) {
    for (const char* n : names)
        if (const char* v = std::getenv(n)) return std::atoi(v);
    return dflt;
}
int launch_rank() { return env_int({"RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID"}, 0); }
int launch_world_size() { return env_int({"WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS"}, 1); }

static bool g_joined = false;

static void send_all(int fd, const void* buf, size_t n) {
    const char* p = static_cast<const char*>(buf);
    while (n > 0) {
        const ssize_t k = ::send(fd, p, n, 0);
        if (k <= 0) cferror("cfMPI_Init: rendezvous send failed");
        p += k; n -= (size_t)k;
    }
}
static void recv_all(int fd, void* buf, size_t n) {
    char* p = static_cast<char*>(buf);
    while (n > 0) {
        const ssize_t k = ::recv(fd, p, n, 0);
        if (k <= 0) cferror("cfMPI_Init: rendezvous receive failed");
        p += k; n -= (size_t)k;
    }
}

int cfMPI_Init(int*, char***) {
    const int rank = launch_rank(), world = launch_world_size();
    if (world <= 1 || g_joined) return 0;
    int crank = 0, cworld = 1;
    cfgpu_check(cfgpu_comm_rank(cfgpu_context(), &crank, &cworld), "cfgpu_comm_rank");
    if (cworld > 1) { g_joined = true; return 0; }  // the host program (or a wrapper) has joined already
    const char* addr = std::getenv("MASTER_ADDR");
    const int port = env_int({"MASTER_PORT"}, 29500) + 1;
    char id[128];
    if (rank == 0) {
        if (cfgpu_comm_unique_id(id) != 0) cferror(std::string("cfMPI_Init: ") + cfgpu_last_error());
        const int ls = ::socket(AF_INET, SOCK_STREAM, 0);
        int one = 1;
        ::setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
        sockaddr_in sa;
        std::memset(&sa, 0, sizeof sa);
        sa.sin_family = AF_INET; sa.sin_addr.s_addr = htonl(INADDR_ANY); sa.sin_port = htons((uint16_t)port);
        if (::bind(ls, reinterpret_cast<sockaddr*>(&sa), sizeof sa) != 0 || ::listen(ls, world) != 0)
            cferror("cfMPI_Init: cannot listen on the rendezvous port " + std::to_string(port));
        for (int k = 1; k < world; ++k) {
            const int fd = ::accept(ls, nullptr, nullptr);
            if (fd < 0) cferror("cfMPI_Init: accept failed");
            send_all(fd, id, sizeof id);
            ::close(fd);
        }
        ::close(ls);
    } else {
        addrinfo hints, *res = nullptr;
        std::memset(&hints, 0, sizeof hints);
        hints.ai_family = AF_INET; hints.ai_socktype = SOCK_STREAM;
        if (::getaddrinfo(addr ? addr : "127.0.0.1", std::to_string(port).c_str(), &hints, &res) != 0 || !res)
            cferror("cfMPI_Init: cannot resolve MASTER_ADDR");
        int fd = -1;
        for (int attempt = 0; attempt < 600; ++attempt) {  // rank 0 may not be listening yet
            fd = ::socket(AF_INET, SOCK_STREAM, 0);
            if (::connect(fd, res->ai_addr, res->ai_addrlen) == 0) break;
            ::close(fd);
            fd = -1;
            ::usleep(100000);
        }
        ::freeaddrinfo(res);
        if (fd < 0) cferror("cfMPI_Init: cannot reach rank 0 for the rendezvous");
        recv_all(fd, id, sizeof id);
        ::close(fd);
    }
    cfgpu_check(cfgpu_comm_init_nccl(cfgpu_context(), rank, world, id), "cfgpu_comm_init_nccl");
    g_joined = true;
    // as the reference: only rank 0 talks (cfmpi.cpp:25-42 mutes cout on the other ranks)
    if (rank != 0) std::cout.setstate(std::ios_base::failbit);
    return 0;
}

int cfMPI_Finalize() {
    std::cout.clear();
    return 0;
}

CfMPI::CfMPI(int nproc0, int nproc1) {
    int r = 0, w = 1;
    cfgpu_check(cfgpu_comm_rank(cfgpu_context(), &r, &w), "cfgpu_comm_rank");
    if (w == 1 && launch_world_size() > 1) {  // constructed before cfMPI_Init: report the launcher's numbers
        r = launch_rank();
        w = launch_world_size();
    }
    taskid_ = r;
    numtasks_ = w;
    usempi_ = w > 1 ? 1 : 0;
    nproc0_ = nproc0 > 0 ? nproc0 : 1;
    nproc1_ = nproc1 > 0 ? nproc1 : (w / nproc0_ > 0 ? w / nproc0_ : 1);
}

}  // namespace chflow
