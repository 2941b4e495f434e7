// rb_common.cuh -- shared declarations of the B200 greedy-sweep engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/rbnics_b200.h"

// ---------------------------------------------------------------------------------------------------------------------
// Tunables (see DESIGN.md "Kernels")
// ---------------------------------------------------------------------------------------------------------------------
constexpr int EST_BM = 128;           // parameters (rows of Z) per CTA
constexpr int EST_BN = 64;            // columns of G' per column tile
constexpr int EST_WARPS = 8;          // 4 (m) x 2 (n), warp tile 32 x 32; no producer warp (last-arriving warp refills)
constexpr int EST_THREADS = EST_WARPS * 32;
constexpr int EST_KSTEP_DOUBLES = 4 * EST_BN;               // one k4-step of one column tile, fragment order (2 KB)
constexpr int EST_NSTAGE = 3;                               // maximum number of G' ring stages (barrier storage)
constexpr int EST_MAX_BLOCKS = 1024;                        // z blocks whose tables are staged in shared memory

constexpr int ASM_BM = 128;  // parameters per CTA of the assembly GEMM
constexpr int ASM_BE = 64;   // matrix entries per CTA

constexpr int LU_MAX_NP_REG = 104;  // largest padded N served by the register-resident LU
constexpr int LU_MAX_NP = 160;      // largest padded N served at all (shared-memory fallback)

// device-resident offline workspace (rb_off_*)
constexpr int RB_OFF_SLOTS = 128;
struct rb_off_csr_t {
  long long* ip = nullptr;
  int* ix = nullptr;
  double* dt = nullptr;
  long long Nh = 0, nnz = 0;
  int sym = 0;  // 1 if the matrix equals its transpose entry by entry (checked on the device at upload)
  long long generation = 0;  // bumped by every rb_off_csr into this slot (invalidates the Gram-Schmidt mirrors)
};
struct rb_off_blk_t {
  double* d = nullptr;  // [Nh][ld] row-major (dof-major)
  long long Nh = 0;
  int cap = 0, n = 0;
  int ld = 0;             // row pitch: cap rounded up to an even count (16-byte aligned rows for the bulk copies)
  // Gram-Schmidt mirrors (allocated by the first rb_off_gram_schmidt on the block): the columns once more, each one
  // contiguous, and X times each column -- the sequential dots / axpys of MGS then read 8 Nh bytes per basis function
  // instead of a 32-byte sector per entry of a dof-major column, and X Z is not recomputed on every call
  double* zt = nullptr;   // [cap][Nh]
  double* xzt = nullptr;  // [cap][Nh]
  int gs_n = 0;           // columns mirrored so far
  int gs_csr = -2;        // operator slot the xzt mirror was built with
  long long gs_csr_generation = -1;  // ... and the generation of that slot
  long long gs_row_begin = -1, gs_row_end = -1;  // dof rows the xzt mirror holds (a rank's partition)
};

// pair-product form of the estimator (rb_est_pair.cu): the quadratic form as one dense GEMM over factor pairs
struct rb_pair_t {
  bool ok = false;                  // the record stream below is valid and rb_estimator_pair_kernel serves the sweeps
  int n_tile = 0, LVP = 0, SK = 0, nstage = 2, vecN = 0;
  long long issued_macs = 0;        // multiply-adds per parameter issued on the tensor pipe
  std::vector<int> h_tile_cost;     // records per column tile
  double* dStream = nullptr;        // records: B fragments of 4 k rows x 64 columns + the factor columns of the 4 rows
  long long* dTileOff = nullptr;    // [n_tile] first record of a tile (offset in doubles)
  int* dTileNrec = nullptr;         // [n_tile]
  int2* dCol = nullptr;             // [n_tile * 64] factor columns (byte offsets in the V' tile) of every column
  // launch plan of the current sweep (prepare_estimator_splits): the row blocks that fill whole waves at the coarse
  // split, and the remaining ones in a second launch with a finer split (shorter CTAs in the last, partial wave)
  int bulk_nsplit = 0, tail_nsplit = 0;
  long long bulk_nrb = 0, tail_nrb = 0;
  int* dSplitBeginTail = nullptr; int tail_cap = 0;
};

// ---------------------------------------------------------------------------------------------------------------------
// Handle
// ---------------------------------------------------------------------------------------------------------------------
struct rb_ctx {
  int device = 0;
  int sm_count = 0, cc_major = 0, cc_minor = 0;
  cudaStream_t stream = nullptr;
  std::string err;

  // reduced system
  bool have_system = false;
  int N = 0, NP = 0, Ql = 0, Qr = 0, QlP = 0, n_bc = 0;
  int E = 0, EP = 0;          // NP*NP and its padding to ASM_BE
  double* dAq = nullptr;      // [QlP][EP] : A_q transposed/padded: entry (i,j) at j*NP+i
  double* dF = nullptr;       // [Qr][NP]
  int* dBcRows = nullptr;     // [n_bc]
  double* dAqRec = nullptr;   // A_q once more as DMMA B-fragment records per 64-entry tile (rb_asm_rec.cu)
  bool asmrec_ok = false;
  int asmrec_KP = 0, asmrec_nslot = 0;

  // estimator
  bool have_est = false;
  int nblk = 0, Kz = 0, Kp = 0, KS = 0, n_ctile = 0, v_off = 0, v_len = 0, LV = 0, LVS = 0, stage_ksteps = 16, est_nstage = 3, est_align = 0, est_tri16 = 0, est_vecN = 0;
  std::vector<int> h_blk_scale, h_blk_src, h_blk_len, h_blk_ks_begin, h_tile_cost, h_tile_b0;
  std::vector<long long> h_tile_off;
  double* dGp = nullptr;            // packed lower-triangular G' in fragment order
  long long gp_doubles = 0;
  long long est_issued_macs = 0;    // multiply-adds per parameter issued by the estimator kernel
  long long* dTileOff = nullptr;    // [n_ctile]
  int* dTileKs0 = nullptr;          // [n_ctile] k-steps skipped per row block of the tile's class
  int* dTileCls = nullptr;          // [n_ctile]
  int* dTileCb = nullptr;           // [4 n_ctile] first row block with data per 16-column range of a tile
  int* dBlkCls = nullptr;           // [nblk] slice class of every z block
  int* dTileB0 = nullptr;           // [n_ctile] block containing the tile's first k-step
  int* dBlkKsBegin = nullptr;       // [nblk+1]
  int* dBlkScale = nullptr;         // [nblk]
  int* dBlkSrc = nullptr;           // [nblk]
  int* dColScale = nullptr;         // [n_ctile*EST_BN]
  int* dColSrc = nullptr;           // [n_ctile*EST_BN]
  // small quadratic forms (Kz <= 128) also keep the dense G and the z map: fused parabolic sweep (rb_parabolic.cu)
  double* dGdense = nullptr;        // G as DMMA B fragments, zero padded to 64 x 64 or 128 x 128 (rb_parabolic_pack_g), or null
  int* dZsrc = nullptr;             // [Kz] index into V
  int* dZscale = nullptr;           // [Kz] theta column or -1
  double* dWeights = nullptr; int weights_cap = 0;  // time-quadrature weights of the parabolic sweep
  rb_pair_t pair;

  // training set
  bool have_train = false;
  int64_t B = 0, Bcap = 0;
  int QT = 0, n_bc_train = 0;
  int64_t idx_off = 0, idx_stride = 1;
  double *dTheta = nullptr, *dThetaBC = nullptr, *dBeta = nullptr;
  // pipelined upload of rb_sweep (host buffers in): Theta rows arrive in pieces on copy_stream, the chunk loop of the
  // sweep waits for the piece it is about to read; up_pieces == 0: everything is resident
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> up_ev;
  long long up_piece = 0;
  int up_pieces = 0;

  // work buffers
  double* dAb = nullptr; int64_t ab_cap = 0;       // assembled matrices of one chunk [chunk][EP]
  double* dU = nullptr; int64_t u_cap = 0; int ldu = 0;
  double* dLuScratch = nullptr; long long lu_scratch_cap = 0;  // L records + packed U of the left-looking LU
  double* dPart = nullptr; int64_t part_cap = 0;   // [nsplit][B]
  double* dDelta = nullptr; double* dEps2 = nullptr; int64_t delta_cap = 0;
  double* dBlkVal = nullptr; long long* dBlkIdx = nullptr; int blk_cap = 0;
  void* dResult = nullptr;           // {double val; int64 idx; int64 n_nonfinite}
  int* dSplitBegin = nullptr; int split_cap = 0;

  rb_off_csr_t off_csr[RB_OFF_SLOTS];
  rb_off_blk_t off_blk[RB_OFF_SLOTS];
  // grow-only scratch of the Gram path: Y = X S2, split partials, C, realigned copies of S and S2
  // (5..8: Gram-Schmidt work vectors)
  void* off_scratch[9] = {};
  size_t off_scratch_cap[9] = {};

  // multi-GPU offline path: in-place SUM all-reduce of a device buffer over the ranks (set by rb_set_allreduce; the
  // host side implements it with NCCL on the very pointer it is handed); nullptr = single rank
  rb_allreduce_fn allreduce = nullptr;
  void* allreduce_user = nullptr;
  long long off_generation = 0;

  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  float last_ms[5] = {0, 0, 0, 0, 0};
  int last_launches = 0;
};

#define RB_CUDA(call)                                                                                       \
  do {                                                                                                      \
    cudaError_t e__ = (call);                                                                               \
    if (e__ != cudaSuccess) {                                                                               \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" +              \
                 std::to_string(__LINE__) + ")";                                                            \
      return RB_ERR_CUDA;                                                                                   \
    }                                                                                                       \
  } while (0)

#define RB_FAIL(code, msg) \
  do {                     \
    ctx->err = (msg);      \
    return (code);         \
  } while (0)

// LU launcher (rb_lu.cu): factor + solve `count` systems assembled in Ab ([count][EP], column-major NP x NP each).
// Theta rows start at theta (already offset to the chunk), U at u.
int rb_launch_lu(rb_ctx* ctx, int64_t count, const double* Ab, const double* theta, const double* thetaBC, double* u);

// pair-product estimator (rb_est_pair.cu); dG = the dense Kz x Kz form on the device
int rb_pair_setup(rb_ctx* ctx, int nblk, const int* scale_norm, const int* blk_src_off, const int* blk_len,
                  const int* cls, const int* ubegin, int v_theta_off, int v_theta_len, int solution_len,
                  const double* dG, int Kz);
int rb_pair_launch(rb_ctx* ctx, const double* Uvec, int ldu, const double* theta, long long rows, int nsplit);
void rb_pair_free(rb_ctx* ctx);

// record-fed assembly GEMM (rb_asm_rec.cu)
int rb_asmrec_setup(rb_ctx* ctx);
int rb_asmrec_launch(rb_ctx* ctx, const double* theta, long long n);
void rb_asmrec_free(rb_ctx* ctx);

// ---------------------------------------------------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// D = A*B (zero C operand): starts an accumulation without a separate clearing pass over the accumulator registers
__device__ __forceinline__ void dmma884_init(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%4};"
      : "=d"(c0), "=d"(c1)
      : "d"(a), "d"(b), "d"(0.0));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk asynchronous copy global -> shared (TMA engine, SASS UBLKCP), completion on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// numpy.argmax ordering: NaN beats everything, then larger value, then lower index.
__device__ __forceinline__ bool argmax_better(double av, long long ai, double bv, long long bi) {
  bool an = (av != av), bn = (bv != bv);
  if (an != bn) return an;
  if (!an && av != bv) return av > bv;
  return ai < bi;
}
#endif
