// eigh_trd.cuh -- interface between the pieces of the tridiagonalisation + divide-and-conquer eigensolver
// (eigh_trd.cu: Householder tridiagonalisation, eigh_dc.cu: D&C on the tridiagonal matrix, back-transformation
// and the driver that kfacb_eigh_batched dispatches to).
#pragma once
#include "common.cuh"

namespace kfacb {
namespace trd {

constexpr int NB = 32;            // panel width = warp size: lane k owns panel column k
constexpr int THREADS = 512;       // 128 registers per thread: 8 x 16-byte loads in flight per lane in the product
constexpr int NWARP = THREADS / 32;
constexpr int MAXN = 8192;        // the reflector and the transposed-product vectors of a column live in shared memory
constexpr int NPART = 66;         // partial sums per team member: [0,32) W^T v, [32,64) V^T v, [64] v.y, [65] |x|^2
constexpr int TS = 160;           // slots per quantity (>= largest team): part[k * TS + rank]

// one matrix to tridiagonalise: T (n x ldt, symmetric copy of the input) -> d, e, tau; reflector j is left in
// T[j][j+1 .. n) (v_j[0] = 1 stored explicitly)
struct Job {
    float* T;
    float* d;
    float* e;
    float* tau;
    float* Vp;            // n x NB panel of reflectors
    float* Wp;            // n x NB panel of W
    float* x;             // n: column being reduced
    float* y;             // n: T v
    float* part;          // NPART x TS partial sums
    float* zpart;         // team x ldt: every CTA's transposed contributions z (symmetric product from the lower triangle)
    unsigned* bar;        // team barrier counter (zero on entry)
    long long* prof;      // NULL, or 9 clock64 totals of the phases (debug: KFACB_TRD_PROF=1)
    const void* tmap;     // CUtensorMap of T (device copy, 128-byte slot): 2-D tiled, box = 128 columns x 8 rows, no swizzle
    int n, ldt, team;
};

struct Sched {            // host-built: CTA c runs entries [begin[c], begin[c+1]) of (job, rank)
    int grid;             // CTAs to launch
};

// Chooses a team size per job and packs the teams on `sms` CTAs (moldable-task list scheduling on a simple
// latency + bandwidth cost model).  Outputs, per job, team[j]; and the per-CTA schedule as flat arrays
// (cta_begin: sms + 1 ints; entries: 2 ints (job, rank) each).  Returns the number of entries.
int plan_teams(const int* n, int njobs, int sms, int* team, int* cta_begin, int* entries);

// launches the persistent kernel (cooperative launch: all CTAs co-resident)
size_t smem_bytes(int n, int team);     // dynamic shared memory a CTA needs for an n x n job in a team of `team`
cudaError_t launch(const Job* d_jobs, const int* d_cta_begin, const int* d_entries, int grid, size_t smem, cudaStream_t st);

// encodes the tensor map of an n x ldt work array at device address T into a 128-byte host slot (the caller uploads it)
int encode_tmap(void* slot128, const float* T, int n, int ldt);

// symmetrised copy (lower triangle mirrored) of the inputs into the work arrays; src[j] n x n, dst = jobs[j].T
cudaError_t launch_copy(const Job* d_jobs, const float* const* d_src, int njobs, int nmax, cudaStream_t st);

}  // namespace trd

// driver of the whole n > 128 path (eigh_dc.cu)
size_t eigh_tdc_workspace_bytes(const int32_t* n, int count, const int* idx, int nb);
int eigh_tdc_run(const int32_t* n, const int* idx, int nb, const float* const* M, float* const* Q, float* const* d,
                 float eps_floor, int32_t* sweeps_out, char* ws, size_t ws_bytes, cudaStream_t st);

int eigh_tdc_last_profile(double* out8);

}  // namespace kfacb
