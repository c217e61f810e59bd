// Multi-GPU through the C ABI with NO Python (and no torch / NCCL) anywhere: what a Rust host would do.
//   cpp_dist_test [P]     P = number of GPUs (default: all visible, at most 8)
// 1. one process per GPU ("rank" mode): the parent forks P children BEFORE touching CUDA; every child calls
//    sfc_comm_init_rank (rendezvous over POSIX shared memory), builds slab plans in both layouts and both directions and
//    checks its share of the distributed fftn against the single-GPU sfc_fftn / sfc_ifftn of the whole volume;
// 2. one process driving all GPUs ("local" mode): sfc_comm_init_local + sfc_dist_exec_host on the whole volume, then
//    sfc_set_num_gpus(P) so that plain sfc_fftn / sfc_execute_batch run over all GPUs.
// Reference seam: trait Communicator + slab partition, scirs2-fft/src/distributed.rs:85-103, 356-362.
// Exit code 0 = pass.  Without a CUDA device: checks the loud BackendError (no CPU fallback).
#include <sys/wait.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <cstring>
#include <vector>

#include "scirs2_fft_cuda.h"

#define REQUIRE(c)                                                                          \
    do {                                                                                    \
        if (!(c)) {                                                                         \
            std::printf("FAILED: %s (line %d): %s\n", #c, __LINE__, sfc_last_error());      \
            return 1;                                                                       \
        }                                                                                   \
    } while (0)

static void fill(std::vector<double>& v, uint64_t seed) {
    uint64_t s = seed * 6364136223846793005ULL + 1442695040888963407ULL;
    for (double& x : v) {
        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
        x = (double)((int64_t)(s >> 11) - (1LL << 52)) / (double)(1LL << 52);
    }
}

static double rel_l2(const double* a, const double* b, size_t n) {
    long double num = 0, den = 0;
    for (size_t i = 0; i < n; ++i) {
        num += (long double)(a[i] - b[i]) * (a[i] - b[i]);
        den += (long double)b[i] * b[i];
    }
    return (double)std::sqrt((double)(num / den));
}

static sfc_dist_desc slab_desc(const int64_t* n, int inverse, int layout) {
    sfc_dist_desc d;
    std::memset(&d, 0, sizeof d);
    d.base.ndim = 3;
    d.base.naxes = 3;
    for (int i = 0; i < 3; ++i) {
        d.base.shape[i] = n[i];
        d.base.axes[i] = i;
    }
    d.base.kind = SFC_C2C;
    d.base.prec = SFC_PREC_F64;
    d.base.direction = inverse ? SFC_INVERSE : SFC_FORWARD;
    d.base.scale = 1.0;
    d.decomposition = SFC_DECOMP_SLAB;
    d.layout = layout;
    return d;
}

// single-GPU reference through the drop-in free functions (norm "forward" on the inverse = unscaled)
static int reference(const std::vector<double>& x, const int64_t* n, int inverse, std::vector<double>& ref) {
    ref.resize(x.size());
    int64_t oshape[3];
    const int64_t total = n[0] * n[1] * n[2];
    return inverse ? sfc_ifftn(x.data(), 3, n, SFC_C128, nullptr, nullptr, 0, "forward", ref.data(), total, oshape)
                   : sfc_fftn(x.data(), 3, n, SFC_C128, nullptr, nullptr, 0, nullptr, ref.data(), total, oshape);
}

static int run_rank(const char* name, int rank, int P) {
    REQUIRE(sfc_init(rank) == SFC_OK);
    sfc_comm* comm = nullptr;
    REQUIRE(sfc_comm_init_rank(&comm, name, rank, P, rank) == SFC_OK);
    REQUIRE(sfc_comm_size(comm) == P && sfc_comm_rank(comm) == rank);
    const int64_t shapes[2][3] = {{64, 64, 64}, {(int64_t)16 * P, (int64_t)8 * P, 64}};
    for (const auto& n : shapes) {
        const int64_t total = n[0] * n[1] * n[2], s0 = n[0] / P, s1 = n[1] / P;
        std::vector<double> x(2 * (size_t)total), ref;
        fill(x, 42 + (uint64_t)n[0]);
        for (int inverse = 0; inverse < 2; ++inverse) {
            REQUIRE(reference(x, n, inverse, ref) == SFC_OK);
            for (int layout = 0; layout < 2; ++layout) {
                sfc_dist_desc d = slab_desc(n, inverse, layout);
                sfc_dist_plan* plan = nullptr;
                REQUIRE(sfc_dist_plan_create(&plan, comm, &d) == SFC_OK);
                sfc_dist_info info;
                REQUIRE(sfc_dist_plan_get_info(plan, &info) == SFC_OK);
                REQUIRE(info.world == P && info.rank == rank && info.local_in_elems == s0 * n[1] * n[2]);
                REQUIRE(info.num_exchanges == (P == 1 ? 0 : (layout == SFC_SLAB_NATURAL ? 2 : 1)));
                const size_t slab = 2 * (size_t)(s0 * n[1] * n[2]);
                std::vector<double> out(slab), want(slab);
                if (layout == SFC_SLAB_NATURAL || P == 1) {
                    std::memcpy(want.data(), ref.data() + (size_t)rank * slab, slab * sizeof(double));
                } else {  // rank r holds out[:, r*s1:(r+1)*s1, :]
                    for (int64_t i = 0; i < n[0]; ++i)
                        std::memcpy(want.data() + 2 * (size_t)(i * s1 * n[2]),
                                    ref.data() + 2 * (size_t)((i * n[1] + rank * s1) * n[2]), 2 * (size_t)(s1 * n[2]) * sizeof(double));
                }
                for (int it = 0; it < 3; ++it) {  // both receive buffers, growing epochs
                    std::fill(out.begin(), out.end(), 0.0);
                    REQUIRE(sfc_dist_exec_host(plan, x.data() + (size_t)rank * slab, out.data()) == SFC_OK);
                    const double e = rel_l2(out.data(), want.data(), slab);
                    if (!(e <= 1e-12)) {
                        std::printf("rank %d shape %lldx%lldx%lld inverse %d layout %d call %d: rel-L2 %.3e\n", rank, (long long)n[0],
                                    (long long)n[1], (long long)n[2], inverse, layout, it, e);
                        return 1;
                    }
                }
                REQUIRE(sfc_dist_plan_destroy(plan) == SFC_OK);
            }
        }
    }
    REQUIRE(sfc_comm_barrier(comm) == SFC_OK);
    REQUIRE(sfc_comm_destroy(comm) == SFC_OK);
    std::printf("rank %d of %d: slab fftn through the C ABI ok\n", rank, P);
    return 0;
}

static int run_local(int P) {
    sfc_comm* comm = nullptr;
    REQUIRE(sfc_comm_init_local(&comm, P, nullptr) == SFC_OK);
    const int64_t n[3] = {64, 128, 64};
    const int64_t total = n[0] * n[1] * n[2];
    std::vector<double> x(2 * (size_t)total), ref, out(2 * (size_t)total);
    fill(x, 7);
    REQUIRE(reference(x, n, 0, ref) == SFC_OK);
    sfc_dist_desc d = slab_desc(n, 0, SFC_SLAB_NATURAL);
    sfc_dist_plan* plan = nullptr;
    REQUIRE(sfc_dist_plan_create(&plan, comm, &d) == SFC_OK);
    for (int it = 0; it < 2; ++it) {
        std::fill(out.begin(), out.end(), 0.0);
        REQUIRE(sfc_dist_exec_host(plan, x.data(), out.data()) == SFC_OK);
        REQUIRE(rel_l2(out.data(), ref.data(), out.size()) <= 1e-12);
    }
    REQUIRE(sfc_dist_plan_destroy(plan) == SFC_OK);
    REQUIRE(sfc_comm_destroy(comm) == SFC_OK);
    // the drop-in free functions over all GPUs of this process
    REQUIRE(sfc_set_num_gpus(P) == SFC_OK && sfc_get_num_gpus() == P);
    int64_t oshape[3];
    std::fill(out.begin(), out.end(), 0.0);
    REQUIRE(sfc_fftn(x.data(), 3, n, SFC_C128, nullptr, nullptr, 0, nullptr, out.data(), total, oshape) == SFC_OK);
    REQUIRE(rel_l2(out.data(), ref.data(), out.size()) <= 1e-12);
    {  // ParallelExecutor::execute_batch (planning_parallel.rs:316-405): rows split over the GPUs, no exchange
        const int64_t count = 4 * P + 1, size = 1000;
        std::vector<double> in(2 * (size_t)(count * size)), a(in.size()), b(in.size());
        fill(in, 9);
        REQUIRE(sfc_execute_batch(in.data(), a.data(), count, size, 0) == SFC_OK);
        REQUIRE(sfc_set_num_gpus(1) == SFC_OK);
        REQUIRE(sfc_execute_batch(in.data(), b.data(), count, size, 0) == SFC_OK);
        REQUIRE(rel_l2(a.data(), b.data(), a.size()) <= 1e-14);
    }
    std::printf("local mode (%d GPUs in one process): slab fftn, sfc_set_num_gpus + sfc_fftn / sfc_execute_batch ok\n", P);
    return 0;
}

int main(int argc, char** argv) {
    // NOTE: no CUDA call before the forks (sfc_device_count would create a context the children inherit broken)
    int P = argc > 1 ? std::atoi(argv[1]) : 0;
    if (argc > 2 && !std::strcmp(argv[2], "--no-device")) {
        sfc_comm* c = nullptr;
        REQUIRE(sfc_comm_init_local(&c, 2, nullptr) == SFC_ERR_BACKEND);
        REQUIRE(sfc_comm_init_rank(&c, "nodev", 0, 1, 0) == SFC_ERR_BACKEND);
        std::puts("no device: BackendError ok");
        return 0;
    }
    if (P <= 0) {
        // count the devices in a child so that the parent stays CUDA-free
        int fd[2];
        REQUIRE(pipe(fd) == 0);
        pid_t pid = fork();
        if (pid == 0) {
            int n = sfc_device_count();
            (void)!write(fd[1], &n, sizeof n);
            _exit(0);
        }
        REQUIRE(read(fd[0], &P, sizeof P) == (ssize_t)sizeof P);
        waitpid(pid, nullptr, 0);
        if (P > 8) P = 8;
        while (P & (P - 1)) --P;  // power of two
    }
    if (P < 1) {
        std::puts("no device");
        return 2;
    }
    char name[64];
    std::snprintf(name, sizeof name, "cpp%d_%ld", (int)getpid(), (long)time(nullptr));
    std::vector<pid_t> kids;
    for (int r = 0; r < P; ++r) {
        pid_t pid = fork();
        if (pid == 0) _exit(run_rank(name, r, P));
        kids.push_back(pid);
    }
    int bad = 0;
    for (pid_t k : kids) {
        int st = 0;
        waitpid(k, &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) ++bad;
    }
    if (bad) {
        std::printf("%d rank(s) failed\n", bad);
        return 1;
    }
    if (run_local(P) != 0) return 1;
    std::puts("cpp dist ok");
    return 0;
}
