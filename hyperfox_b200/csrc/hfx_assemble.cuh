// Fused per-element HDG kernel for sm_100a: geometry -> operator contractions -> structured static condensation ->
// Dirichlet masking -> deterministic scatter into the global trace matrix (block CSR), one thread group per element (32..256 threads
// by element size), everything in shared memory, the results leave as bulk asynchronous copies.
//
// Replaces, per element, the reference's HDGSolver::calcElementalMatrices (src/solver/HDGSolver.cpp:176-359) with
// Model::compute (src/model/HDG*.cpp), the operators (src/operator/HDGBase.cpp:67-158, HDGDiffusion.cpp:74-145,
// HDGConvection.cpp:60-104, Reaction.cpp, Source.cpp, Euler.cpp:28-30), applyBoundaryConditions (:361-529, CGType models)
// and assembleSystem (:531-675).
//
// Structure exploited (valid for every in-scope model, checked numerically against the oracle in tests/):
//   S_qq = M (x) I_dim        (HDGBase.cpp:152; no other operator writes the q rows)      -> one nN x nN inverse W = M^-1
//   q-rows / l-rows couple to u,q only through the face nodes                              -> t x t weighted face mass matrices
//
// Shared-memory layout rule: matrices read as the LEFT operand of a product are column-major (row index contiguous), matrices
// read as the RIGHT operand are row-major (column index contiguous); leading dimensions are even so that every register tile
// is fetched with 128-bit shared loads.
#pragma once
#include <cstdint>
#include <cstdlib>

namespace hfx {

struct AsmParams {
  int nCells;
  int eBegin, eEnd;           // element range of this launch (a pipelined assemble launches one range per upload piece)
  // mesh
  const double* elemX;        // [nCells][nN*DIM] element-major node coordinates (gathered once at allocate)
  const int* cells; const int* cell2face;
  // per-element maps built on device by elem_maps_kernel (hfx_allocate)
  const uint8_t* fperm;       // [nCells][nFc*nNf] position in faces[F] of element-local face node (HDGSolver.cpp:258-275)
  const uint8_t* tauSide;     // [nCells][nFc]     0 if this cell is face2Cell[F][0] (HDGSolver.cpp:290-293)
  const uint8_t* elemPos;     // [nCells][nFc*nFc] position of face f2 in the sorted neighbour list of face f
  // per-face
  const long long* faceRowStart;  // [nFaces] offset of row (F,0) in vals
  const uint8_t* faceNnb;         // [nFaces] number of neighbour faces (row length = nnb*t)
  const uint8_t* faceBC;          // [nFaces] 0 none, 1 Dirichlet, 2 integrated Dirichlet
  const uint8_t* faceInterior;    // [nFaces] 1 if two adjacent cells
  // fields
  const double* tau; int tauVals;
  const double* diff; int diffComps; int diffIsCell;
  double diffConst;           // D = diffConst * I when there is no diffusion field (1, or the value of a scalar DiffusionTensor that is constant over the mesh)
  const double* vel;
  const double* srcIP; const double* reacIP;
  const double* solOld;
  const double* dirichlet;
  int opmask; int timeScheme; double dt;
  // tables (device global, read-only)
  const double* shape; const double* dshape; const double* w;
  const double* fshape; const double* fdshape; const double* fw;
  const double* ffs;          // [nIPf][nNf*nNf] phi_a*phi_b products of the face element
  const int* faceNodes;       // [nFc][nNf]
  const int8_t* nodeInFace;   // [nFc][nN] inverse of faceNodes (-1 if not on the face)
  const double* mhinv;        // inverse of the reference mass matrix, column-major [ev(nN)][ev(nN)] (unit pad diagonal)
  // straight-sided elements: reference matrices of the purely geometric blocks (built by hfx_refel_set) and the per-element flag
  const double* sref;         // S^_r [DIM][nN][ev(nN)]: Squ_d = sum_r detJ Jinv(d,r) S^_r
  const double* eref;         // E_f [nFc][nN][ev(nN)]: reference face mass scattered to the element nodes, entry (f, j, i) = M^f[a_f(i)][a_f(j)] or 0
  const double* srefT;        // S^_r transposed [DIM][nN][ev(nN)]: entry (r, j, k) = S^_r[k][j], the layout of the Suq_d left operand
  int noRef;                  // experiments: disable the all-reference path
  const double* aref;         // A^_r [DIM][ev(nN) x nN] column-major: A_d = sum_r Jinv(d,r) A^_r          (A^_r = M_ref^-1 S^_r)
  const double* mfref;        // M^f [ev(t) x t]: face reference mass
  const double* bref;         // B^_f [nFc][nN][t] = M_ref^-1[:, faceNodes_f] M^f: W Sql_d = -(area n_d / detJ) B^_f
  const uint8_t* affine;      // [nCells] or NULL (shortcut disabled)
  const double* colTab;       // tables of the column-per-lane kernel (hfx_col.cuh) or NULL
  // outputs
  double* U; double* Q; double* U0; double* Q0; double* S; double* S0;  // S,S0 may be NULL
  double* vals; double* rhs;
  int* status;                // bit 0: (near) zero pivot met in a local inverse
  long long* prof;            // optional [16] per-phase cycle counters (block 0 only), NULL in production
  // recovery by recomputation (hfx_allocate flag HFX_RECOMPUTE_RECOVERY: U, Q are not stored; hfx_recover re-condenses the element and applies
  // u_e = U lambda_e + U0, q_e = Q lambda_e + Q0 (HDGSolver.cpp:741-775) out of shared memory).  Served by hfx_big.cuh.
  int gjThreads;              // hfx_big.cuh: threads of the K^-1 Gauss-Jordan (experiments: HFX_BIG_GJ)
  int recover; const double* recTrace; double* recSol; double* recFlux;
};

template <int DIM, int P> struct ElemCfg;
#define HFX_CFG(D, PP, NN, NNF, NIP, NIPF) \
  template <> struct ElemCfg<D, PP> { static constexpr int nN = NN, nNf = NNF, nIP = NIP, nIPf = NIPF, nFc = D + 1; };
// nN, nNf, nIP (degree 2p), nIPf   (Cubature.cpp nIP map)
HFX_CFG(2, 1, 3, 2, 3, 2)
HFX_CFG(2, 2, 6, 3, 6, 3)
HFX_CFG(2, 3, 10, 4, 12, 4)
HFX_CFG(2, 4, 15, 5, 16, 5)
HFX_CFG(2, 5, 21, 6, 25, 6)
HFX_CFG(3, 1, 4, 3, 4, 3)
HFX_CFG(3, 2, 10, 6, 14, 6)
HFX_CFG(3, 3, 20, 10, 24, 12)
HFX_CFG(3, 4, 35, 15, 46, 16)
HFX_CFG(3, 5, 56, 21, 81, 25)
#undef HFX_CFG

constexpr int kAsmThreads = 256;
#define HFX_PROF(i) do { if (p.prof && blockIdx.x == 0 && threadIdx.x == 0) { long long c_ = clock64(); p.prof[i] += c_ - tprev; tprev = c_; } } while (0)

__host__ __device__ constexpr int ev(int x) { return (x + 1) & ~1; }

// ---- register-tile micro kernel ------------------------------------------------------------------------------------------
// acc[i][j] += A[i] * B[j] over one k: A column-major (16-byte aligned, TM even), B row-major (TN even) => 128-bit shared loads.
template <int TM, int TN>
__device__ __forceinline__ void mk_step(double (&acc)[TM][TN], const double* __restrict__ a, const double* __restrict__ b) {
  double av[TM], bv[TN];
#pragma unroll
  for (int i = 0; i < TM; i += 2) { const double2 v = *reinterpret_cast<const double2*>(a + i); av[i] = v.x; av[i + 1] = v.y; }
#pragma unroll
  for (int j = 0; j < TN; j += 2) { const double2 v = *reinterpret_cast<const double2*>(b + j); bv[j] = v.x; bv[j + 1] = v.y; }
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
}
template <int TM, int TN, int K>
__device__ __forceinline__ void mk(double (&acc)[TM][TN], const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb) {
  if (K <= 24) {
#pragma unroll
    for (int k = 0; k < K; k++) mk_step<TM, TN>(acc, A + lda * k, B + ldb * k);
  } else {
#pragma unroll 4
    for (int k = 0; k < K; k++) mk_step<TM, TN>(acc, A + lda * k, B + ldb * k);
  }
}
template <int TM, int TN>
__device__ __forceinline__ void zero_acc(double (&acc)[TM][TN]) {
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.0;
}

// ---- Gauss-Jordan inverse by one warp --------------------------------------------------------------------------------------
// Lane r owns rows r, r+32, ... in registers; the pivot row of each step is published in a scratch line in shared memory and read
// back as a broadcast, so no shuffles.  Unpivoted (M and K are definite for a coercive local problem); a vanishing pivot raises
// bit 0 of *flag.  src/dst are column-major with leading dimension ld (src may equal dst).
template <int n>
__device__ __forceinline__ void warp_invert(const double* src, double* dst, int ld, double* scratch /*[2*ev(n)]*/, int lane, int* flag) {
  constexpr int R = (n + 31) / 32, np = ev(n);
  double row[R][n];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int i = lane + 32 * r;
#pragma unroll
    for (int j = 0; j < n; j++) row[r][j] = (i < n) ? src[i + ld * j] : 0.0;
  }
  double scale = 0.0;
#pragma unroll
  for (int r = 0; r < R; r++)
#pragma unroll
    for (int j = 0; j < n; j++) scale = fmax(scale, fabs(row[r][j]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) scale = fmax(scale, __shfl_xor_sync(0xffffffffu, scale, o));
  bool bad = false;
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < n; j++) scratch[j] = row[0][j];
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < n; k++) {
    const double* pr = scratch + (k & 1) * np;     // pivot row before the update
    double* nx = scratch + ((k + 1) & 1) * np;
    const double piv = pr[k];
    if (!(fabs(piv) > 1e-14 * scale)) bad = true;
    const double ip = 1.0 / piv;
#pragma unroll
    for (int r = 0; r < R; r++) {
      const bool isPivotRow = (lane + 32 * r) == k;
      const double f = row[r][k] * ip;
#pragma unroll
      for (int j = 0; j < n; j++) {
        const double pj = pr[j];
        if (j == k) row[r][j] = isPivotRow ? ip : -f;
        else row[r][j] = isPivotRow ? pj * ip : fma(-f, pj, row[r][j]);
      }
      if (k + 1 < n && (lane + 32 * r) == k + 1) {
#pragma unroll
        for (int j = 0; j < n; j++) nx[j] = row[r][j];
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int i = lane + 32 * r;
    if (i < n) {
#pragma unroll
      for (int j = 0; j < n; j++) dst[i + ld * j] = row[r][j];
    }
  }
  if (bad && lane == 0) atomicOr(flag, 1);
}

// ---- small device helpers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double fast_rcp(double x) {   // MUFU seed + two Newton steps: full double precision, no slow path
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  return r;
}
__device__ __forceinline__ double fast_rsqrt(double x) {   // MUFU seed (2^-22) + two Newton steps: 1/sqrt(x) to ~1 ulp without the sqrt + rcp chain
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double h = 0.5 * x;
  r = fma(fma(-h * r, r, 0.5), r, r);
  r = fma(fma(-h * r, r, 0.5), r, r);
  return r;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
// bulk asynchronous copies shared -> global (TMA, SASS UBLKCP / UBLKRED): sizes and both addresses are multiples of 16 bytes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(double* gdst, const double* ssrc, int bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(sa), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_add_f64(double* gdst, const double* ssrc, int bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(gdst), "r"(sa), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// bulk asynchronous copies global -> shared (TMA, SASS UBLKCP.S.G) completing on an mbarrier
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, int bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
  asm volatile("{\n\t.reg .pred p;\n\tMBAR_WAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra MBAR_WAIT_%=;\n\t}"
               ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// warp-granular dynamic tile queue: every call hands the warp the next 32 consecutive tile ids
__device__ __forceinline__ int grab32(int* ctr, int lane) {
  int v = 0;
  if (lane == 0) v = atomicAdd(ctr, 32);
  return __shfl_sync(0xffffffffu, v, 0) + lane;
}

// ---- FP64 tensor-core building blocks (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4) --------------------------------------------
// Measured on B200 (tools/ubench/dmma.cu): 16 cycles per DMMA per SM sub-partition (same 37 TFLOP/s peak as the DFMA pipe, which it
// shares), latency 26 cycles.  What it buys here is operand traffic: one 64-bit shared load per lane feeds 8 FMAs per lane, where a
// 2x2 CUDA-core register tile needs 8 bytes of shared-memory traffic per FMA and saturates the 128 B/clk shared pipe at 25% of peak.
// Fragment layout: lane = 4*lr + lc;  A[m0+lr][k0+lc], B[k0+lc][n0+lr], C[m0+lr][n0+2*lc+{0,1}].
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
// One warp task: rows [8*mt, 8*mt+8) x NTW column tiles starting at column tile nt0, K = 4*KS.
// fa(m, k), fb(k, n) return operand entries (they own all range handling: k beyond the true K must give an exact zero product);
// fs(m, n, v0, v1) receives C[m][n], C[m][n+1] (n even).
template <int NTW, int KS, class FA, class FB, class FS>
__device__ __forceinline__ void mma_task(int mt, int nt0, int lane, FA fa, FB fb, FS fs) {
  const int lr = lane >> 2, lc = lane & 3;
  const int m = mt * 8 + lr;
  double c[NTW][2];
#pragma unroll
  for (int j = 0; j < NTW; j++) { c[j][0] = 0.0; c[j][1] = 0.0; }
#pragma unroll
  for (int ks = 0; ks < KS; ks++) {
    const int k = ks * 4 + lc;
    const double a = fa(m, k);
#pragma unroll
    for (int j = 0; j < NTW; j++) dmma(c[j], a, fb(k, (nt0 + j) * 8 + lr));
  }
#pragma unroll
  for (int j = 0; j < NTW; j++) fs(m, (nt0 + j) * 8 + 2 * lc, c[j][0], c[j][1]);
}
// single output tile with the K loop split over two accumulators (dependent DMMA chains would otherwise be latency bound)
template <int KS, class FA, class FB, class FS>
__device__ __forceinline__ void mma_task_splitk(int mt, int nt, int lane, FA fa, FB fb, FS fs) {
  const int lr = lane >> 2, lc = lane & 3;
  const int m = mt * 8 + lr, n = nt * 8 + lr;
  double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
#pragma unroll
  for (int ks = 0; ks < KS; ks += 2) {
    const int k = ks * 4 + lc;
    dmma(c0, fa(m, k), fb(k, n));
    if (ks + 1 < KS) dmma(c1, fa(m, k + 4), fb(k + 4, n));
  }
  fs(m, nt * 8 + 2 * lc, c0[0] + c1[0], c0[1] + c1[1]);
}
// Lean variant for operands that are affine in k (pointer + stride): no index arithmetic or range predicates in the k loop.
// pa -> A[m][0] with stride lda between consecutive k; pb[j] -> B[0][n_j] with stride ldb.  Rows / columns beyond the matrix
// are clamped by the caller (their results are discarded); k beyond the true K contributes an exact zero (a = 0, b finite).
template <int NTW, int K>
__device__ __forceinline__ void mma_affine(double (&c)[NTW][2], const double* __restrict__ pa, int lda,
                                           const double* const (&pb)[NTW], int ldb, int lc) {
  constexpr int KS = (K + 3) / 4;
#pragma unroll
  for (int ks = 0; ks < KS; ks++) {
    const int k = ks * 4 + lc;
    double a, b[NTW];
    if (ks * 4 + 3 < K) {
      a = pa[k * lda];
#pragma unroll
      for (int j = 0; j < NTW; j++) b[j] = pb[j][k * ldb];
    } else {
      const int kk = k < K ? k : K - 1;
      a = k < K ? pa[kk * lda] : 0.0;
#pragma unroll
      for (int j = 0; j < NTW; j++) b[j] = pb[j][kk * ldb];
    }
#pragma unroll
    for (int j = 0; j < NTW; j++) dmma(c[j], a, b[j]);
  }
}
template <int NTW>
__device__ __forceinline__ void zero_c(double (&c)[NTW][2]) {
#pragma unroll
  for (int j = 0; j < NTW; j++) { c[j][0] = 0.0; c[j][1] = 0.0; }
}
__device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }

__device__ __forceinline__ int grab1(int* ctr, int lane) {   // warp-granular dynamic task queue
  int v = 0;
  if (lane == 0) v = atomicAdd(ctr, 1);
  return __shfl_sync(0xffffffffu, v, 0);
}

// ---- Gauss-Jordan inverse by a group of kGJThreads threads (warps 0..3, named barrier 1) ---------------------------------
// 2x2 block pivots: half the serial depth of the scalar algorithm (the pivot chain, not the flops, is what costs: ~300-450
// cycles per barrier step on B200, see tools/ubench/gj2.cu).  Ping-pong between two column-major buffers with even leading
// dimension ld: a step reads one buffer and writes the other, so one barrier per pivot block suffices.  Each thread owns 2x1
// strips aligned with the pivot pairs; branch-free.  Unpivoted (M and K are definite for a coercive local problem); a vanishing
// pivot block raises bit 0 of *flag.  np = n rounded up to even: the caller provides pad row/column = 0, pad diagonal = 1.
// The inverse ends in ((np/2) odd ? b1 : b0).
constexpr int kGJThreads = 128;
template <int np, int ld, int GT, bool kWholeCTA = (GT == kAsmThreads)>
__device__ __noinline__ void group_invert(double* b0, double* b1, int tid, int* flag, int barId) {
  // not inlined on purpose: inside the big kernel the register allocator rematerialises every address of this latency-bound loop
  constexpr int MT = np / 2, NS = MT * np, NQ = (NS + GT - 1) / GT;
  const double s0 = b0[0];
  const double scale = s0 * s0;
  bool bad = false;
  // per-thread strip offsets, computed once: strip (rows i0, i0+1 ; column j)
  int offA[NQ], offI[NQ], offJ[NQ], jq[NQ];
#pragma unroll
  for (int q = 0; q < NQ; q++) {
    const int s2 = tid + q * GT, sc = s2 < NS ? s2 : 0;
    const int i0 = (sc % MT) * 2, j = sc / MT;
    offI[q] = i0; offJ[q] = ld * j; offA[q] = i0 + ld * j; jq[q] = s2 < NS ? j : -8;
  }
  // one step: reads src, writes dst (distinct buffers), one barrier.  The update is written with the adjugate, P^-1 = adj(P) / det: every
  // product is formed while the reciprocal of det is still in flight and 1/det enters in the last FMA only.
  auto step = [&](const double* __restrict__ src, double* __restrict__ dst, int k) {
    const double* colk = src + ld * k;
    const double2 pc0 = *reinterpret_cast<const double2*>(colk + k);          // pivot block, column 0
    const double2 pc1 = *reinterpret_cast<const double2*>(colk + ld + k);     // pivot block, column 1
    double2 a[NQ], c0[NQ], c1[NQ], pj[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) {
      a[q] = *reinterpret_cast<const double2*>(src + offA[q]);
      c0[q] = *reinterpret_cast<const double2*>(colk + offI[q]);
      c1[q] = *reinterpret_cast<const double2*>(colk + ld + offI[q]);
      pj[q] = *reinterpret_cast<const double2*>(src + offJ[q] + k);
    }
    const double det = fma(pc0.x, pc1.y, -pc1.x * pc0.y);
    if (!(fabs(det) > 1e-28 * scale)) bad = true;
    const double id = fast_rcp(det);
    const double j00 = pc1.y, j01 = -pc1.x, j10 = -pc0.y, j11 = pc0.x;   // adj(P)
#pragma unroll
    for (int q = 0; q < NQ; q++) {
      if (jq[q] >= 0) {
        const int j = jq[q];
        const bool inK = (j == k) || (j == k + 1);
        double v0 = fma(j00, pj[q].x, j01 * pj[q].y), v1 = fma(j10, pj[q].x, j11 * pj[q].y);   // adj(P) A[K,j]
        if (j == k) { v0 = j00; v1 = j10; }
        if (j == k + 1) { v0 = j01; v1 = j11; }
        const double x0 = fma(c0[q].x, v0, c1[q].x * v1), x1 = fma(c0[q].y, v0, c1[q].y * v1);
        const double ax = inK ? 0.0 : a[q].x, ay = inK ? 0.0 : a[q].y;
        double r0 = fma(-x0, id, ax), r1 = fma(-x1, id, ay);
        if (offI[q] == k) { r0 = v0 * id; r1 = v1 * id; }
        *reinterpret_cast<double2*>(dst + offA[q]) = make_double2(r0, r1);
      }
    }
    if (kWholeCTA) __syncthreads(); else if (GT == 32) __syncwarp(); else bar_sync_named(barId, GT);
  };
#pragma unroll 1
  for (int k = 0; k + 2 < np; k += 4) { step(b0, b1, k); step(b1, b0, k + 2); }
  if ((np / 2) & 1) step(b0, b1, np - 2);
  if (bad) atomicOr(flag, 1);
}

// ---- Gauss-Jordan inverse with 4x4 block pivots and tensor-core updates ----------------------------------------------------------------
// In-place-form block Gauss-Jordan, ping-pong b0 -> b1 -> b0 ... (column-major, leading dimension ld, np a multiple of 4).  Step s, pivot rows /
// columns K = [4s, 4s+4):  T = P^-1 src'[K, :],  dst = base - A' T  with the substitutions that make every entry of the inverse-so-far fall out of
// the same product:  src'[K, K] = I,  A'[K, :] = -I (so the pivot rows become T),  A'[i, :] = src[i, K],  base = src outside the pivot rows / columns, 0 inside.
// P^-1 comes from the adjugate: lane (lr, lc) forms the cofactor it needs as its DMMA operand, det(P) is a quad reduction, ONE reciprocal per four
// pivots (the serial pivot chain, not the flops, is what an inverse costs: 18 barrier steps of ~700 cycles for the 2x2-block version at np = 36).
// Warp w < ceil(np/8) owns the columns [8w, 8w+8) of every step: 1 + ceil(np/8) DMMAs per step.  Unpivoted, like group_invert; the callers refine U.
// The inverse ends in ((np/4) odd ? b1 : b0).  Must be called by whole warps; warps >= ceil(np/8) return at once (the caller synchronises the CTA).
template <int np, int ld>
__device__ __noinline__ void block4_invert(double* b0, double* b1, int tid, int* flag, int barId) {
  static_assert(np % 4 == 0, "4x4 pivot blocks");
  constexpr int NS = np / 4, NCT = (np + 7) / 8;
  const int warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
  if (warp >= NCT) return;
  const double* src = b0; double* dst = b1;
  const double s00 = b0[0];
  const double scale = (s00 * s00) * (s00 * s00);
  bool bad = false;
  // cofactor geometry of this lane: entry (r, c) = (lr & 3, lc) of P^-1 is (-1)^(r+c) det(P without row c and column r) / det P
  const int r = lr & 3, c = lc;
  const int R0 = c == 0 ? 1 : 0, R1 = c <= 1 ? 2 : 1, R2 = c <= 2 ? 3 : 2;      // rows of the minor
  const int C0 = r == 0 ? 1 : 0, C1 = r <= 1 ? 2 : 1, C2 = r <= 2 ? 3 : 2;      // columns of the minor
  const double sgn = ((r + c) & 1) ? -1.0 : 1.0;
  const int n0 = warp * 8;
  const int jb = imin(n0 + lr, np - 1);            // column this lane fetches for T (B fragment)
#pragma unroll 1
  for (int s = 0; s < NS; s++) {
    const int k = 4 * s;
    const double* P = src + k + ld * k;
    const double m00 = P[R0 + ld * C0], m01 = P[R0 + ld * C1], m02 = P[R0 + ld * C2];
    const double m10 = P[R1 + ld * C0], m11 = P[R1 + ld * C1], m12 = P[R1 + ld * C2];
    const double m20 = P[R2 + ld * C0], m21 = P[R2 + ld * C1], m22 = P[R2 + ld * C2];
    // det P from the 2x2 minors of its two row pairs (every lane, no shuffles: the chain is six dependent FP64 operations, and a dependent FP64
    // operation costs ~40 cycles on this part)
    double2 q0[4], q1[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { q0[j] = *reinterpret_cast<const double2*>(P + ld * j); q1[j] = *reinterpret_cast<const double2*>(P + ld * j + 2); }
    // rows 0,1: a0j = q0[j].x, a1j = q0[j].y ; rows 2,3: a2j = q1[j].x, a3j = q1[j].y
    const double s0 = fma(q0[0].x, q0[1].y, -q0[0].y * q0[1].x), s1 = fma(q0[0].x, q0[2].y, -q0[0].y * q0[2].x), s2 = fma(q0[0].x, q0[3].y, -q0[0].y * q0[3].x);
    const double s3 = fma(q0[1].x, q0[2].y, -q0[1].y * q0[2].x), s4 = fma(q0[1].x, q0[3].y, -q0[1].y * q0[3].x), s5 = fma(q0[2].x, q0[3].y, -q0[2].y * q0[3].x);
    const double c5 = fma(q1[2].x, q1[3].y, -q1[2].y * q1[3].x), c4 = fma(q1[1].x, q1[3].y, -q1[1].y * q1[3].x), c3 = fma(q1[1].x, q1[2].y, -q1[1].y * q1[2].x);
    const double c2 = fma(q1[0].x, q1[3].y, -q1[0].y * q1[3].x), c1 = fma(q1[0].x, q1[2].y, -q1[0].y * q1[2].x), c0 = fma(q1[0].x, q1[1].y, -q1[0].y * q1[1].x);
    const double det = fma(s0, c5, fma(-s1, c4, s2 * c3)) + fma(s3, c2, fma(-s4, c1, s5 * c0));
    // operands that do not depend on P^-1 are fetched while the cofactor chain runs
    const bool jInK = (unsigned)(jb - k) < 4u;
    const double tsrc = jInK ? ((jb - k) == lc ? 1.0 : 0.0) : src[(k + lc) + ld * jb];
    const double d0 = fma(m11, m22, -m12 * m21), d1 = fma(m10, m22, -m12 * m20), d2 = fma(m10, m21, -m11 * m20);
    const double adj = sgn * fma(m00, d0, fma(-m01, d1, m02 * d2));
    if (!(fabs(det) > 1e-56 * scale)) bad = true;
    // one Newton step on the reciprocal seed (relative error ~1e-12): the callers refine U = -K^-1 R with one step of iterative refinement, which
    // squares what an inexact K^-1 leaves behind
    double rd;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rd) : "d"(det));
    rd = fma(fma(-det, rd, 1.0), rd, rd);
    const double pinv = lr < 4 ? adj * rd : 0.0;
    double t[2] = {0.0, 0.0};
    dmma(t, pinv, tsrc);                           // T[lr][n0 + 2 lc + {0,1}], rows lr < 4
    // T as the right operand of the update: B[kk = lc][n = lr] = T[lc][n0 + lr], held by lane 4 lc + (lr >> 1), component lr & 1
    const int sl = 4 * lc + (lr >> 1);
    const double t0 = __shfl_sync(0xffffffffu, t[0], sl), t1 = __shfl_sync(0xffffffffu, t[1], sl);
    const double bt = (lr & 1) ? t1 : t0;
#pragma unroll
    for (int mt = 0; mt < NCT; mt++) {
      const int i = mt * 8 + lr, ic = imin(i, np - 1);
      const bool iInK = (unsigned)(i - k) < 4u;
      const double a = iInK ? ((i - k) == lc ? 1.0 : 0.0) : -src[ic + ld * (k + lc)];      // -A'
      double cc[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int j = n0 + 2 * lc + h, jc = imin(j, np - 1);
        cc[h] = (iInK || (unsigned)(j - k) < 4u) ? 0.0 : src[ic + ld * jc];
      }
      dmma(cc, a, bt);
      if (i < np) {
#pragma unroll
        for (int h = 0; h < 2; h++) { const int j = n0 + 2 * lc + h; if (j < np) dst[i + ld * j] = cc[h]; }
      }
    }
    bar_sync_named(barId, NCT * 32);
    const double* tsw = dst; dst = const_cast<double*>(src); src = tsw;
  }
  if (bad) atomicOr(flag, 1);
}

__device__ __forceinline__ double det_only(const double (&J)[2][2]) { return J[0][0] * J[1][1] - J[0][1] * J[1][0]; }
__device__ __forceinline__ double det_only(const double (&J)[3][3]) {
  const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2], c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  return J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
}
__device__ __forceinline__ void det_inv(const double (&J)[2][2], double& det, double (&I)[2][2]) {
  det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
  const double id = fast_rcp(det);
  I[0][0] = J[1][1] * id; I[0][1] = -J[0][1] * id; I[1][0] = -J[1][0] * id; I[1][1] = J[0][0] * id;
}
__device__ __forceinline__ void det_inv(const double (&J)[3][3], double& det, double (&I)[3][3]) {
  const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2], c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
  const double id = fast_rcp(det);
  I[0][0] = c00 * id; I[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id; I[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
  I[1][0] = c01 * id; I[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id; I[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
  I[2][0] = c02 * id; I[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id; I[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
}

template <int DIM, int P>
struct AsmSmem {
  using C = ElemCfg<DIM, P>;
  static constexpr int nN = C::nN, t = C::nNf, nFc = C::nFc, nIP = C::nIP, nIPf = C::nIPf;
  static constexpr int l = nFc * t, NW = 3 + 2 * DIM;   // face weight kinds: tau, n_d, (Dn)_d, v.n, 1
  static constexpr int nNp = ev(nN), tp = ev(t), ldc = ev(l + 1) + 2, nJ = nIP + nFc * nIPf;
  static constexpr int ldg = ev(DIM * nN);              // leading dimension of the g rows (one row per ip)
  static constexpr int ldw = ((nFc * NW + 3) / 4) * 4;
  // offsets in doubles (all even => 16-byte aligned)
  static constexpr int oPHI = 0;                             // shape table, row-major [nIP][nNp]
  static constexpr int oWQ = oPHI + nIP * nNp;               // cubature weights [nIP] then face weights [nIPf] (resident)
  static constexpr int oX = oWQ + ev(nIP + nIPf);            // coords [nN][DIM]
  static constexpr int oIJ = oX + ev(nN * DIM);              // dV * invJ  [nIP][DIM*DIM]  (m,r)
  static constexpr int oDV = oIJ + ev(nIP * DIM * DIM);      // dV [nIP]
  static constexpr int oDIP = oDV + ev(nIP);                 // D at bulk IPs [nIP][DIM*DIM] col-major
  static constexpr int oVIP = oDIP + ev(nIP * DIM * DIM);    // v at bulk IPs [nIP][DIM]
  static constexpr int oLW = oVIP + ev(nIP * DIM);           // (reac*dV + euler*dV) [nIP], src*dV [nIP]
  static constexpr int oFWT = oLW + ev(2 * nIP);             // face IP weights, row-major [nIPf][ldw], column (f,kind)
  static constexpr int oTAU = oFWT + nIPf * ldw;             // tau at element-local face nodes [l]
  static constexpr int oDN = oTAU + ev(l);                   // D at nodes [nN][DIM*DIM]
  static constexpr int oVN = oDN + ev(nN * DIM * DIM);       // v at nodes [nN][DIM]
  static constexpr int oG = oVN + ev(nN * DIM);              // g rows [nIP][ldg], entry (d,i) -> later A_d [DIM][nNp x nN] col-major
  static constexpr int szG0 = (nIP * ldg > DIM * nNp * nN) ? nIP * ldg : DIM * nNp * nN;
  static constexpr int szG = (szG0 > nJ * DIM * DIM) ? szG0 : nJ * DIM * DIM;   // also hosts the raw Jacobians [nJ][DIM*DIM] (dead before g is formed)
  static constexpr int oJ = oG;
  static constexpr int oM = oG + ev(szG) + 4;                // M, col-major ld nNp
  static constexpr int oW = oM + nNp * nNp;                  // W = M^-1 (nNp columns: the block Gauss-Jordan works on the even-padded matrix)
  static constexpr int oSQU = oW + nNp * nNp;                // Squ_d row-major [DIM][nN][nNp] -> later U row-major [nN][ldc]
  static constexpr int szSQU = (DIM * nN * nNp > nN * ldc) ? DIM * nN * nNp : nN * ldc;
  static constexpr int oSUQ = oSQU + szSQU;                  // Suq_d col-major [DIM][nNp x nN]
  static constexpr int oSUU = oSUQ + DIM * nNp * nN;         // Suu -> K -> K^-1, col-major ld nNp
  static constexpr int oFW = oSUU + nNp * nNp;               // weighted face mass matrices [nFc][NW][tp x t] (symmetric)
  static constexpr int oB = oFW + nFc * NW * tp * t;         // B_d row-major [DIM][nN][ldc] -> Q_d
  static constexpr int oR = oB + DIM * nN * ldc;             // R row-major [nN][ldc]; before P6 it hosts the suu left operand cg [nIP][nNp]
  static constexpr int szR = (nN * ldc > nIP * nNp) ? nN * ldc : nIP * nNp;
  static constexpr int oCG = oR;
  // Per-element staging of the read-only tables (with ~227 KB of shared memory carved out the L1 is ~1 KB, so __ldg would go to L2):
  //   dshape | fdshape | fshape at the start of the FW region (all three are consumed before the face matrices are formed),
  //   ffs (phi_a phi_b of the face element) at the end of the B region (consumed while FW is written, before B is).
  static constexpr int nDSH = ev(nIP * nN * DIM), nFDS = ev(nIPf * t * (DIM - 1)), nFSH = ev(nIPf * t), nFFS = ev(nIPf * t * t);
  static constexpr int oDSH = oFW, oFDS = oDSH + nDSH, oFSH = oFDS + nFDS, oFFS = oR - nFFS;
  static_assert(oFSH + nFSH <= oFFS, "staged tables do not fit in the FW+B regions");
  static_assert(oFFS >= oB, "ffs staging must not overlap the face matrices");
  static constexpr int oFU = oR + szR;                       // Fu [nN]
  static constexpr int oSCR = oFU + ev(nN) + 2;              // (FU[ev(nN)] holds 1/detJ of the first cubature point)
  static constexpr int oGEO = oSCR + 2 * nNp;                // constant geometry of a straight-sided element: Jinv [DIM*DIM], det, per face n[DIM], area
  static constexpr int oGEOR = oGEO + ev(DIM * DIM + 1 + nFc * (DIM + 1));   // all-reference path, per face: -area n_d [DIM], tau area, area
  static constexpr int oEnd = oGEOR + ev(nFc * (DIM + 2));
  static_assert(ev(nFc * nN * t) <= szR && tp * t <= nNp * nNp, "reference tables are staged in the R and M regions");
  static_assert(nFc * nN * nNp <= nNp * nNp + szSQU, "node-scattered face masses are staged in the W + Squ regions");
  // S staging ([l][ldc], or the 16 t x t blocks of the bulk write-out): reuses the dead g/A + M + W span when it is large enough (large elements),
  // otherwise gets its own area (small elements, where shared memory is not the limit)
  static constexpr bool stFits = (oSQU - oG) >= l * ldc;
  static constexpr int oST = stFits ? oG : oEnd;
  static constexpr int nDoubles = stFits ? oEnd : oEnd + l * ldc;
  // after the doubles: row starts (nFc x int64) then a small int area
  static constexpr int nNLUT = ((DIM * t + 7) / 8) * 8;              // column (d,b) of B_d -> offset of row block d + b
  static constexpr int nKLUT = (((1 + DIM) * t + 3) / 4) * 4;        // P9 reduction index (kind,b) -> operand offsets
  static constexpr int nInts = 8 + 2 * nFc * t + nFc * nN + nFc * nFc + 5 * nFc + 8 + nNLUT + nFc * nKLUT + nFc * l;   // (KLUT: two 16-bit offsets per entry)
  static_assert(nDoubles < 65536, "16-bit operand offsets");
  static constexpr size_t bytes = (size_t)nDoubles * 8 + 8 * (nFc + l + 1) + 4 * (size_t)nInts;   // (+1: the mbarrier of the bulk loads)
  static constexpr size_t gbytes = (bytes + 15) & ~(size_t)15;   // per element group
};

// TPE = threads per element: the CTA's 256 threads form 256 / TPE independent groups, each with its own shared-memory image and its
// own element stream (group-local barriers only).  One CTA-wide group for p=3 tets, one warp per element for the linear elements.
template <int DIM, int P, int TPE>
__global__ void __launch_bounds__(kAsmThreads, 2) hdg_assemble_kernel(const AsmParams p) {
  using L = AsmSmem<DIM, P>;
  constexpr int nN = L::nN, t = L::t, nFc = L::nFc, nIP = L::nIP, nIPf = L::nIPf, l = L::l, NW = L::NW;
  constexpr int nNp = L::nNp, tp = L::tp, ldc = L::ldc, ldg = L::ldg, ldw = L::ldw, nJ = L::nJ;
  constexpr int D2 = DIM * DIM, NT = TPE, NWARP = NT / 32, NGRP = kAsmThreads / TPE;
  static_assert(TPE == 32 || TPE == 64 || TPE == 128 || TPE == 256, "group size");
  static_assert(l < NT && nFc * nFc <= NT, "one thread per trace row");
  constexpr int kTau = 0, kN = 1, kDN = 1 + DIM, kC = 1 + 2 * DIM, kOne = 2 + 2 * DIM;
  constexpr int FWS = tp * t;   // stride between face matrices
  constexpr bool kBulkUQ = (l % 2) == 0;   // rows of U, Q are multiples of 16 bytes: bulk copies
  constexpr bool kBulkS = (t % 2) == 0;    // (row, face block) pieces of S are multiples of 16 bytes: bulk copies / reduce-adds
  constexpr bool kPrefetch = (TPE == 256) && (nN * DIM <= 64) && (l <= 128) && (nFc * nFc <= 32);
  extern __shared__ __align__(16) double sm_all[];
  const int grp = threadIdx.x / TPE;
  double* const sm = sm_all + (size_t)grp * (L::gbytes / 8);
  auto gsync = [&]() { if (TPE == kAsmThreads) __syncthreads(); else if (TPE == 32) __syncwarp(); else bar_sync_named(1 + grp, TPE); };
  auto gsync_or = [&](int pred) -> int {
    if (TPE == kAsmThreads) return __syncthreads_or(pred);
    if (TPE == 32) { __syncwarp(); return __any_sync(0xffffffffu, pred); }
    unsigned r;
    asm volatile("{ .reg .pred p, q; setp.ne.s32 p, %1, 0; barrier.cta.red.or.pred q, %2, %3, p; selp.u32 %0, 1, 0, q; }" : "=r"(r) : "r"(pred), "r"(1 + grp), "r"(TPE) : "memory");
    return (int)r;
  };
  double* PHI = sm + L::oPHI; double* X = sm + L::oX; double* JR = sm + L::oJ; double* IJ = sm + L::oIJ; double* DV = sm + L::oDV;
  double* DIP = sm + L::oDIP; double* VIP = sm + L::oVIP; double* LW = sm + L::oLW; double* FWT = sm + L::oFWT; double* TAU = sm + L::oTAU;
  double* DN = sm + L::oDN; double* VN = sm + L::oVN; double* G = sm + L::oG; double* CG = sm + L::oCG; double* Mm = sm + L::oM;
  double* Wb = sm + L::oW; double* SQU = sm + L::oSQU; double* SUQ = sm + L::oSUQ; double* SUU = sm + L::oSUU; double* FW = sm + L::oFW;
  double* B = sm + L::oB; double* R = sm + L::oR; double* FU = sm + L::oFU; double* GEO = sm + L::oGEO; double* GEOR = sm + L::oGEOR;
  double* WQ = sm + L::oWQ; double* DSH = sm + L::oDSH; double* FDS = sm + L::oFDS; double* FSH = sm + L::oFSH; double* FFS = sm + L::oFFS;
  long long* ROWS = reinterpret_cast<long long*>(sm + L::nDoubles);   // [nFc] first entry of row (F,0) in vals
  long long* RBASE = ROWS + nFc;                                      // [l] first entry of the CSR row of element-local trace row r
  unsigned long long* MBAR = reinterpret_cast<unsigned long long*>(RBASE + l);   // completion barrier of the bulk loads of the reference matrices
  int* ISM = reinterpret_cast<int*>(RBASE + l + 1);                   // [nFc] global face ids
  int* FN = ISM + 8;                                                  // [nFc*t] faceNodes
  int* PERM = FN + nFc * t;                                           // [nFc*t] element-local -> face-node position
  int* NIF = PERM + nFc * t;                                          // [nFc*nN] node -> position in face (or -1)
  int* POS = NIF + nFc * nN;                                          // [nFc*nFc]
  int* RLEN = POS + nFc * nFc;                                        // [nFc] row length
  int* BCF = RLEN + nFc;                                              // [nFc] boundary kind
  int* INTF = BCF + nFc;                                              // [nFc] interior flag
  int* OPP = INTF + nFc;                                              // [nFc] first node not on the face (orientation test)
  int* QCTR = OPP + nFc;                                              // [4] dynamic tile-queue counters
  int* NLUT = QCTR + 8;                                               // [nNLUT]
  unsigned short* KLUT = reinterpret_cast<unsigned short*>(NLUT + L::nNLUT);   // [nFc][nKLUT][2] 16-bit operand offsets of the P9 reduction index k = (kind, b)
  int* CMAP = NLUT + L::nNLUT + nFc * L::nKLUT;                                    // (bulk write-out) [l+1], shares the POSROW area
  int* POSROW = NLUT + L::nNLUT + nFc * L::nKLUT;                                  // [nFc][l] column offset of element-local column cc inside a row of face f
  double* A = G;    // A_d aliases g (dead after the contractions)
  double* ST = sm + L::oST;   // S staging for the write-out
  double* Um = SQU; // U aliases Squ (dead after A)
  const int tid = threadIdx.x & (TPE - 1), lane = tid & 31, warp = tid >> 5;
  const bool hasDiff = p.opmask & 1, hasConv = p.opmask & 2, hasReac = (p.opmask & 4) && p.reacIP, hasSrc = (p.opmask & 8) && p.srcIP;
  const bool euler = p.timeScheme == 1;
  const double ts = euler ? p.dt : 1.0;   // Euler::apply scales the u rows (Su, Fu) by dt before adding the mass terms (Euler.cpp:28-32)
  const bool diffField = hasDiff && p.diffComps > 0 && p.diff;
  const double dsc = hasDiff ? (diffField ? 1.0 : p.diffConst) : 0.0;   // scale of the D = c I blocks (Suq, Slq)
  const bool needSuu = hasConv || hasReac || euler;   // bulk part of Suu: -C^T, reaction mass, Euler mass
  // All-reference path: a straight-sided element of a Laplace-type model (no convection / reaction / time scheme / diffusion field)
  // whose tau is constant on each face has EVERY block of its local matrix as a scalar combination of reference matrices
  // (tau mass = tau_f area_f M^f as well), so no cubature-point work is left at all.
  const bool refModel = !diffField && !needSuu && !p.noRef && p.srefT;
  constexpr bool kRefSrcPf = kPrefetch && nIP <= 64;

  // ---- once per CTA: constant tables; padding lanes must hold finite numbers --------------------------------------------
  for (int i = tid; i < L::nDoubles; i += NT) sm[i] = 0.0;
  if (tid == 0) { mbar_init(MBAR, 1); fence_proxy_async(); }
  gsync();
  for (int i = tid; i < nFc * t; i += NT) FN[i] = p.faceNodes[i];
  for (int i = tid; i < nFc * nN; i += NT) NIF[i] = p.nodeInFace[i];
  for (int i = tid; i < nIP * nN; i += NT) PHI[(i / nN) * nNp + (i % nN)] = p.shape[i];
  for (int i = tid; i < nIP + nIPf; i += NT) WQ[i] = i < nIP ? p.w[i] : p.fw[i - nIP];
  if (tid < nFc) { int vn = 0; for (int kk = 0; kk < nN; kk++) if (p.nodeInFace[tid * nN + kk] < 0) { vn = kk; break; } OPP[tid] = vn; }
  for (int n = tid; n < L::nNLUT; n += NT) { const int nn = n < DIM * t ? n : DIM * t - 1; NLUT[n] = (nn / t) * nN * ldc + (nn % t); }
  gsync();

  // reference-mass inverse: each thread keeps its share in registers for the whole kernel (constant-detJ shortcut for M^-1)
  constexpr int NMH = (nNp * nNp + NT - 1) / NT;
  constexpr bool kMHReg = NMH <= 2;
  double mh[kMHReg ? NMH : 1];
  if (kMHReg) {
#pragma unroll
    for (int q = 0; q < NMH; q++) mh[q] = (tid + q * NT < nNp * nNp) ? p.mhinv[tid + q * NT] : 0.0;
  }

  // ---- software prefetch of the next element's gather (registers) ------------------------------------------------------
  double pfX = 0.0, pfTau = 0.0, pfSrc = 0.0;
  int pfF = 0, pfPerm = 0, pfPos = 0, pfRlen = 0, pfBc = 0, pfInt = 0;
  long long pfRow = 0;
  // two stages: the second one's addresses depend on values loaded by the first (face ids), so it is issued a few phases later --
  // issuing both back to back parks the issuing warps on the first loads' DRAM latency in the middle of the geometry phase
  int pfSide = 0, pfAff = 0;
  auto prefetchA = [&](int e) {
    if (!kPrefetch || e >= p.eEnd) return;
    pfAff = p.affine ? p.affine[e] : 0;   // every thread: warp-uniform control flow later, no shared-memory round trip
    if (tid < nN * DIM) pfX = p.elemX[(size_t)e * nN * DIM + tid];
    if (tid >= 64 && tid < 64 + l) {
      const int i = tid - 64, f = i / t;
      pfF = p.cell2face[(size_t)e * nFc + f];
      pfPerm = p.fperm[(size_t)e * l + i];
      pfSide = (p.tauVals == 2) ? p.tauSide[(size_t)e * nFc + f] : 0;
    }
    if (tid >= 192 && tid < 192 + nFc) pfF = p.cell2face[(size_t)e * nFc + (tid - 192)];
    if (tid >= 224 && tid < 224 + nFc * nFc) pfPos = p.elemPos[(size_t)e * nFc * nFc + (tid - 224)];
    if (kRefSrcPf && refModel && hasSrc && tid >= 128 && tid < 128 + nIP) pfSrc = p.srcIP[(size_t)e * nIP + (tid - 128)];
  };
  auto prefetchB = [&](int e) {
    if (!kPrefetch || e >= p.eEnd) return;
    if (tid >= 64 && tid < 64 + l) pfTau = p.tau[((size_t)pfF * t + pfPerm) * p.tauVals + pfSide];
    if (tid >= 192 && tid < 192 + nFc) { const int F = pfF; pfRow = p.faceRowStart[F]; pfRlen = (int)p.faceNnb[F] * t; pfBc = p.faceBC[F]; pfInt = p.faceInterior[F]; }
  };
  const int e0 = p.eBegin + blockIdx.x * NGRP + grp, eStride = gridDim.x * NGRP;
  prefetchA(e0);
  prefetchB(e0);

  // column tiles (8 columns, kind-major) of the weighted face mass contraction that this model needs
  constexpr int NCT = (nFc * NW + 7) / 8;
  unsigned baseNeed = 0, intNeed = 0, affNeed = 0;   // affNeed: straight-sided elements take the n_d kinds from the reference face mass
#pragma unroll
  for (int T = 0; T < NCT; T++) {
#pragma unroll
    for (int cidx = 8 * T; cidx < 8 * T + 8; cidx++) {
      if (cidx < nFc * NW) {
        const int kind = cidx / nFc;
        if (kind <= DIM || (kind < kC && diffField) || (kind == kC && hasConv)) baseNeed |= 1u << T;
        if (kind == 0 || (kind > DIM && kind < kC && diffField) || (kind == kC && hasConv)) affNeed |= 1u << T;
        if (kind == kOne) intNeed |= 1u << T;
      }
    }
  }
  const int kDNe = diffField ? kDN : kN;   // D = I (HDGDiffusion.cpp:102-105): (Dn)_d mass = n_d mass
  // P9 reduction index k = (kind, b): kind 0 -> Slu (tau mass on U), kind 1+d -> Slq_d ((Dn)_d mass, stored negated, on Q_d):
  // offset of the left operand inside the face's weighted mass matrices, offset of the right operand's row (U or Q_d, node fn_f(b))
  for (int i = tid; i < nFc * L::nKLUT; i += NT) {
    const int f = i / L::nKLUT, k = i - f * L::nKLUT, kk = k < (1 + DIM) * t ? k : (1 + DIM) * t - 1;
    const int kind = kk / t, b = kk - kind * t, nd = FN[f * t + b];
    KLUT[2 * i] = (unsigned short)((kind == 0 ? kTau : kDNe + kind - 1) * FWS + tp * b);
    KLUT[2 * i + 1] = (unsigned short)(kind == 0 ? L::oSQU + nd * ldc : L::oB + ((kind - 1) * nN + nd) * ldc);
  }
  gsync();

  unsigned mphase = 0;
  long long tprev = clock64();
  for (int e = e0; e < p.eEnd; e += eStride) {
    // ---- P0: commit the prefetched gather -----------------------------------------------------------------------------------
    if (kPrefetch) {
      if (tid < nN * DIM) X[tid] = pfX;
      if (tid >= 64 && tid < 64 + l) { TAU[tid - 64] = pfTau; PERM[tid - 64] = pfPerm; }
      if (tid >= 192 && tid < 192 + nFc) { const int f = tid - 192; ISM[f] = pfF; ROWS[f] = pfRow; RLEN[f] = pfRlen; BCF[f] = pfBc; INTF[f] = pfInt; }
      if (tid >= 224 && tid < 224 + nFc * nFc) POS[tid - 224] = pfPos;
    } else {   // large elements: direct gather
      for (int i = tid; i < nN * DIM; i += NT) X[i] = p.elemX[(size_t)e * nN * DIM + i];
      for (int i = tid; i < l; i += NT) {
        const int f = i / t;
        const int F = p.cell2face[(size_t)e * nFc + f];
        const int pos = p.fperm[(size_t)e * l + i];
        PERM[i] = pos;
        const int side = (p.tauVals == 2) ? p.tauSide[(size_t)e * nFc + f] : 0;
        TAU[i] = p.tau[((size_t)F * t + pos) * p.tauVals + side];
      }
      if (tid < nFc) {
        const int F = p.cell2face[(size_t)e * nFc + tid];
        ISM[tid] = F; ROWS[tid] = p.faceRowStart[F]; RLEN[tid] = (int)p.faceNnb[F] * t; BCF[tid] = p.faceBC[F]; INTF[tid] = p.faceInterior[F];
      }
      for (int i = tid; i < nFc * nFc; i += NT) POS[i] = p.elemPos[(size_t)e * nFc * nFc + i];
    }
    if (tid == NT - 1) { QCTR[0] = 0; QCTR[1] = 0; QCTR[2] = 0; QCTR[3] = 0; }
    // straight-sided element (flag computed once at allocate): constant Jacobians straight from the vertices, and the purely geometric
    // blocks Squ, A = Sqq^-1 Squ, B = Sqq^-1 Sql and the n_d face matrices are scalar combinations of reference matrices
    const bool aff = (kPrefetch ? pfAff : (p.affine ? (int)p.affine[e] : 0)) != 0;
    const bool affAB = aff && !diffField;              // with a diffusion field g (which A aliases) is still live when the tables are applied
    const bool needG = !aff || hasConv || diffField;   // the gradient rows g are only needed by the convection / diffusion-field contractions
    const bool refCand = aff && refModel;              // all-reference path if, in addition, tau is constant on each face (checked below)
    // stage the read-only tables of this element pass (L2 -> shared, asynchronous 16-byte copies)
    auto stageStd = [&]() {
      if (needG) for (int i = tid; i < L::nDSH / 2; i += NT) cp_async16(DSH + 2 * i, p.dshape + 2 * i);
      if (!aff) for (int i = tid; i < L::nFDS / 2; i += NT) cp_async16(FDS + 2 * i, p.fdshape + 2 * i);
      for (int i = tid; i < L::nFSH / 2; i += NT) cp_async16(FSH + 2 * i, p.fshape + 2 * i);
      for (int i = tid; i < L::nFFS / 2; i += NT) cp_async16(FFS + 2 * i, p.ffs + 2 * i);
    };
    if (refCand) {
      // reference matrices, staged where their combinations end up (A^_r -> A_d and S^_r^T -> Suq_d are combined in place) or in
      // regions that are idle until the condensation (B^_f in R, M^f in M)
      // five bulk copies (TMA) issued by one thread, completing on the group's mbarrier: 39 KB at p=3 without a single per-thread copy instruction
      if (tid == 0) {
        constexpr int szA = DIM * nN * nNp * 8, szB = ev(nFc * nN * t) * 8, szM = FWS * 8, szE = nFc * nN * nNp * 8;
        mbar_expect_tx(MBAR, 2 * szA + szB + szM + szE);
        bulk_load(A, p.aref, szA, MBAR); bulk_load(SUQ, p.srefT, szA, MBAR);
        bulk_load(R, p.bref, szB, MBAR); bulk_load(Mm, p.mfref, szM, MBAR);
        bulk_load(Wb, p.eref, szE, MBAR);   // E_f spans the W and Squ regions
      }
      if (hasSrc) {   // source values times the cubature weights (scaled by det J later)
        if (kRefSrcPf) { if (tid >= 128 && tid < 128 + nIP) LW[nIP + tid - 128] = ts * pfSrc * WQ[tid - 128]; }
        else for (int i = tid; i < nIP; i += NT) LW[nIP + i] = ts * p.srcIP[(size_t)e * nIP + i] * WQ[i];
      }
    } else {
      stageStd();
      if (affAB) {   // straight-sided element of any other model: the geometric blocks still come from staged reference matrices (B^_f in the idle Squ region)
        for (int i = tid; i < DIM * nN * nNp / 2; i += NT) { cp_async16(A + 2 * i, p.aref + 2 * i); cp_async16(SUQ + 2 * i, p.srefT + 2 * i); }
        for (int i = tid; i < ev(nFc * nN * t) / 2; i += NT) cp_async16(SQU + 2 * i, p.bref + 2 * i);
        for (int i = tid; i < FWS / 2; i += NT) cp_async16(Mm + 2 * i, p.mfref + 2 * i);
      }
    }
    const double* const brefS = refCand ? R : SQU;   // where the staged B^_f sits
    const int* cell = p.cells + (size_t)e * nN;
    if (diffField) {
      for (int i = tid; i < nN * D2; i += NT) {
        const int nd = i / D2, c = i % D2;
        const size_t ent = p.diffIsCell ? ((size_t)e * nN + nd) : (size_t)cell[nd];
        double v;
        if (p.diffComps == 1) v = ((c / DIM) == (c % DIM)) ? p.diff[ent] : 0.0;
        else v = p.diff[ent * D2 + c];
        DN[i] = v;
      }
    }
    if (hasConv) for (int i = tid; i < nN * DIM; i += NT) VN[i] = p.vel[(size_t)cell[i / DIM] * DIM + (i % DIM)];
    if (!refCand) cp_async_wait_all();
    gsync();
    HFX_PROF(0);
    unsigned tileNeed = aff ? affNeed : baseNeed;
#pragma unroll
    for (int f = 0; f < nFc; f++) if (BCF[f] == 2) tileNeed |= intNeed;
    // scatter maps of this element (HDGSolver.cpp:596: matRowCols), consumed by the write-out
    // Global storage: per face F its nnb(F) neighbour blocks, each a contiguous row-major t x t block in face-node order (block CSR).
    if (kBulkS) {   // staging already is in face-node order: CMAP[position] = element-local trace index
      if (tid < l) CMAP[(tid / t) * t + PERM[tid]] = tid;
      if (tid == l) CMAP[l] = l;   // (l < NT)
    } else {
      for (int i = tid; i < nFc * l; i += NT) { const int f = i / l, cc = i - f * l; POSROW[i] = POS[f * nFc + cc / t] * t * t + PERM[cc]; }
      if (tid < l) { const int f = tid / t; RBASE[tid] = ROWS[f] + (long long)PERM[tid] * t; }
    }

    // ---- PG (all-reference path, stage G): constant geometry (one thread per face + one for the bulk Jacobian), tau check ----------
    bool ref = false;
    if (refCand) {
      int bad = 0;
      if (tid < nFc) {
        const int f = tid;
        const int* fn = FN + f * t;
        double J[DIM - 1][DIM];
#pragma unroll
        for (int r = 0; r < DIM - 1; r++)
#pragma unroll
          for (int m = 0; m < DIM; m++) J[r][m] = 0.5 * (X[fn[r + 1] * DIM + m] - X[fn[0] * DIM + m]);
        double nv[DIM], area;
        if (DIM == 2) {
          nv[0] = -J[0][1]; nv[1] = J[0][0];
        } else {
          nv[0] = J[0][1] * J[DIM - 2][2 % DIM] - J[0][2 % DIM] * J[DIM - 2][1];
          nv[1] = J[0][2 % DIM] * J[DIM - 2][0] - J[0][0] * J[DIM - 2][2 % DIM];
          nv[DIM - 1] = J[0][0] * J[DIM - 2][1] - J[0][1] * J[DIM - 2][0];
        }
        // |J_0 x J_1| = sqrt(det(J J^T)) (Operator.cpp:66-69; 2-D: |tangent|): the norm of the un-normalised normal is the face measure
        double nn = 0.0;
#pragma unroll
        for (int m = 0; m < DIM; m++) nn = fma(nv[m], nv[m], nn);
        const double inrm = fast_rsqrt(nn);
        area = nn * inrm;
        // outward orientation (HDGBase.cpp:43-62) from the un-normalised normal, beside the rsqrt chain; area * n = the oriented un-normalised normal
        const int v0 = fn[0], vn = OPP[f];
        double prod = 0.0;
#pragma unroll
        for (int m = 0; m < DIM; m++) prod = fma(X[vn * DIM + m] - X[v0 * DIM + m], nv[m], prod);
        const double sg = prod > 0.0 ? -1.0 : 1.0;
        double* gf = GEO + D2 + 1 + f * (DIM + 1);
        double* gr = GEOR + f * (DIM + 2);
#pragma unroll
        for (int m = 0; m < DIM; m++) { gr[m] = -sg * nv[m]; gf[m] = sg * nv[m] * inrm; }
        gf[DIM] = area;
        gr[DIM] = TAU[f * t] * area;
        gr[DIM + 1] = area;
      } else if (tid == (NT > 32 ? 32 : nFc)) {
        double J[DIM][DIM], det, I[DIM][DIM];
#pragma unroll
        for (int r = 0; r < DIM; r++)
#pragma unroll
          for (int m = 0; m < DIM; m++) J[r][m] = 0.5 * (X[(r + 1) * DIM + m] - X[m]);
        det_inv(J, det, I);
#pragma unroll
        for (int m = 0; m < DIM; m++)
#pragma unroll
          for (int r = 0; r < DIM; r++) GEO[m * DIM + r] = I[m][r];
        GEO[D2] = det;
        FU[ev(nN)] = fast_rcp(det);
      } else if (NT > 64 && tid >= 64) {
        for (int i = tid - 64; i < l; i += NT - 64) bad |= (TAU[i] != TAU[(i / t) * t]);
      }
      if (NT <= 64) for (int i = tid; i < l; i += NT) bad |= (TAU[i] != TAU[(i / t) * t]);
      mbar_wait(MBAR, mphase); mphase ^= 1;   // the staged reference matrices have landed
      ref = !gsync_or(bad);
      HFX_PROF(5);
      prefetchA(e + eStride);
      if (!ref) {   // tau varies on a face: the general straight-sided path needs its own tables
        stageStd();
        cp_async_wait_all();
        gsync();
      }
    }
    if (ref) {
      // ---- PR (all-reference path, stage R): every block from the staged reference matrices ---------------------------------------------------------------
      //   A_d = sum_r Jinv(d,r) A^_r ;  Suq_d = detJ sum_r Jinv(d,r) S^_r^T - (area n_d) face mass ;  Suu = tau area face mass
      //   B_d = -(area n_d / detJ) B^_f ;  face matrices tau area M^f, -area n_d M^f, area M^f
      const double det = GEO[D2], rdet = FU[ev(nN)];
      double Ii[DIM][DIM];
#pragma unroll
      for (int m = 0; m < DIM; m++)
#pragma unroll
        for (int r = 0; r < DIM; r++) Ii[m][r] = GEO[m * DIM + r];
      // One list of independent work items spread round-robin over the threads (16-byte shared-memory accesses):
      //   [A] A_d pairs, [S] Suq_d / Suu pairs (face parts from the node-scattered face masses E_f), [B] B_d, [F] face matrices.
      constexpr int NA2 = nN * nNp / 2, NF2 = nFc * FWS / 2;
      constexpr bool kVecT = (t % 2) == 0;
      constexpr int NB = kVecT ? nFc * nN * t / 2 : nFc * nN * t;
      constexpr int I_S = NA2, I_B = 2 * NA2, I_F = I_B + NB, I_END = I_F + NF2;
      double2* const A2 = reinterpret_cast<double2*>(A);
      double2* const SUQ2 = reinterpret_cast<double2*>(SUQ);
      const double2* const E2 = reinterpret_cast<const double2*>(Wb);
      for (int item = tid; item < I_END; item += NT) {
        if (item < I_S) {
          const int idx2 = item;
          double2 ar[DIM];
#pragma unroll
          for (int r = 0; r < DIM; r++) ar[r] = A2[r * NA2 + idx2];
#pragma unroll
          for (int d = 0; d < DIM; d++) {
            double2 va = make_double2(0.0, 0.0);
#pragma unroll
            for (int r = 0; r < DIM; r++) { va.x = fma(Ii[d][r], ar[r].x, va.x); va.y = fma(Ii[d][r], ar[r].y, va.y); }
            A2[d * NA2 + idx2] = va;
          }
        } else if (item < I_B) {
          const int idx2 = item - I_S;
          double2 sr[DIM], fq[DIM], suu = make_double2(0.0, 0.0);
#pragma unroll
          for (int r = 0; r < DIM; r++) { sr[r] = SUQ2[r * NA2 + idx2]; fq[r] = make_double2(0.0, 0.0); }
#pragma unroll
          for (int f = 0; f < nFc; f++) {
            const double2 ev2 = E2[f * NA2 + idx2];
            const double* gr = GEOR + f * (DIM + 2);
            suu.x = fma(gr[DIM], ev2.x, suu.x); suu.y = fma(gr[DIM], ev2.y, suu.y);
#pragma unroll
            for (int d = 0; d < DIM; d++) { fq[d].x = fma(gr[d], ev2.x, fq[d].x); fq[d].y = fma(gr[d], ev2.y, fq[d].y); }
          }
          *reinterpret_cast<double2*>(SUU + 2 * idx2) = make_double2(ts * suu.x, ts * suu.y);   // index i + nNp j = 2 idx2 (pad rows: zeros)
#pragma unroll
          for (int d = 0; d < DIM; d++) {
            double2 vs = make_double2(0.0, 0.0);
#pragma unroll
            for (int r = 0; r < DIM; r++) { vs.x = fma(Ii[d][r], sr[r].x, vs.x); vs.y = fma(Ii[d][r], sr[r].y, vs.y); }
            SUQ2[d * NA2 + idx2] = make_double2(ts * dsc * fma(det, vs.x, fq[d].x), ts * dsc * fma(det, vs.y, fq[d].y));
          }
        } else if (item < I_F) {
          const int ib = item - I_B;
          if (kVecT) {
            constexpr int t2 = kVecT ? t / 2 : 1;
            const int f = ib / (nN * t2), rem = ib - f * nN * t2, m = rem / t2, b = 2 * (rem - m * t2);
            const double* gr = GEOR + f * (DIM + 2);
            const double2 rv = reinterpret_cast<const double2*>(R)[ib];
            const double s0 = rdet * rv.x, s1 = rdet * rv.y;
#pragma unroll
            for (int d = 0; d < DIM; d++) *reinterpret_cast<double2*>(B + (d * nN + m) * ldc + f * t + b) = make_double2(s0 * gr[d], s1 * gr[d]);
          } else {
            const int f = ib / (nN * t), rem = ib - f * nN * t, m = rem / t, b = rem - m * t;
            const double* gr = GEOR + f * (DIM + 2);
            const double sc = rdet * R[ib];
#pragma unroll
            for (int d = 0; d < DIM; d++) B[(d * nN + m) * ldc + f * t + b] = sc * gr[d];
          }
        } else {
          const int i2 = item - I_F, f = i2 / (FWS / 2), ab2 = i2 - f * (FWS / 2);
          const double2 mv = reinterpret_cast<const double2*>(Mm)[ab2];
          const double* gr = GEOR + f * (DIM + 2);
          double2* fw2 = reinterpret_cast<double2*>(FW + f * NW * FWS) + ab2;
          fw2[kTau * FWS / 2] = make_double2(gr[DIM] * mv.x, gr[DIM] * mv.y);
#pragma unroll
          for (int d = 0; d < DIM; d++) fw2[(kN + d) * FWS / 2] = make_double2(gr[d] * mv.x, gr[d] * mv.y);
          fw2[kOne * FWS / 2] = make_double2(gr[DIM + 1] * mv.x, gr[DIM + 1] * mv.y);
        }
      }
      if ((nN & 1) && tid == NT - 1) {   // odd size: unit pad diagonal of K for the 2x2-block Gauss-Jordan (the pad row already holds zeros)
        for (int j = 0; j < nN; j++) SUU[j + nNp * nN] = 0.0;
        SUU[nN + nNp * nN] = 1.0;
      }
      for (int idx = tid; idx < DIM * nN; idx += NT) { B[idx * ldc + l] = 0.0; B[idx * ldc + l + 1] = 0.0; }  // Q0 column
      if (tid >= NT - 32) {   // Fu = source (Source.cpp:24-48)
        for (int i = tid - (NT - 32); i < nN; i += 32) {
          double s2 = 0.0;
          if (hasSrc) for (int ip = 0; ip < nIP; ip++) s2 = fma(PHI[ip * nNp + i], LW[nIP + ip], s2);
          FU[i] = s2 * det;
        }
      }
      gsync();
      HFX_PROF(1);
      prefetchB(e + eStride);
    }
    double* const W = ((nNp / 2) & 1) ? Wb : Mm;
    constexpr int MTN = (nN + 7) / 8;           // 8-row tiles over the element nodes
    constexpr int KS_IP = (nIP + 3) / 4, KS_N = (nN + 3) / 4;
    if (!ref) {
    // ---- P1a: raw Jacobians (Operator.cpp:14-39) as two small tensor-core products: rows (ip, r), reduction over the nodes,
    //      columns = the DIM coordinates (one 8-wide tile, DIM columns used) ------------------------------------------------------
    if (!aff) {
      const int lr = lane >> 2, lc = lane & 3;
      constexpr int MB = nIP * DIM, MBT = (MB + 7) / 8, MF = nIPf * (DIM - 1), MFT = (MF + 7) / 8;
      constexpr int KSB = (nN + 3) / 4, KSF1 = (t + 3) / 4;
      const int mcol = imin(lr, DIM - 1);   // B-operand column held by this lane
      for (int task = warp; task < MBT + nFc * MFT; task += NWARP) {
        double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
        if (task < MBT) {
          const int row = task * 8 + lr, rc = imin(row, MB - 1);
          const double* pa = DSH + (rc / DIM) * nN * DIM + (rc % DIM);     // dshape[ip][i][r], stride DIM in i
#pragma unroll
          for (int ks = 0; ks < KSB; ks++) {
            const int k = ks * 4 + lc, kk = imin(k, nN - 1);
            const double a = k < nN ? pa[kk * DIM] : 0.0;
            const double b = X[kk * DIM + mcol];
            if (ks & 1) dmma(c1, a, b); else dmma(c0, a, b);
          }
          if (row < MB) {
            if (2 * lc < DIM) JR[row * DIM + 2 * lc] = c0[0] + c1[0];
            if (2 * lc + 1 < DIM) JR[row * DIM + 2 * lc + 1] = c0[1] + c1[1];
          }
        } else {
          const int tf = task - MBT, f = tf / MFT, row = (tf % MFT) * 8 + lr, rc = imin(row, MF - 1);
          const int* fn = FN + f * t;
          const double* pa = FDS + (rc / (DIM - 1)) * t * (DIM - 1) + (rc % (DIM - 1));   // fdshape[ip][a][r], stride DIM-1 in a
#pragma unroll
          for (int ks = 0; ks < KSF1; ks++) {
            const int k = ks * 4 + lc, kk = imin(k, t - 1);
            const double a = k < t ? pa[kk * (DIM - 1)] : 0.0;
            const double b = X[fn[kk] * DIM + mcol];
            if (ks & 1) dmma(c1, a, b); else dmma(c0, a, b);
          }
          if (row < MF) {
            double* dst = JR + (nIP + f * nIPf + row / (DIM - 1)) * D2 + (row % (DIM - 1)) * DIM;
            if (2 * lc < DIM) dst[2 * lc] = c0[0] + c1[0];
            if (2 * lc + 1 < DIM) dst[2 * lc + 1] = c0[1] + c1[1];
          }
        }
      }
    }
    if (!aff) gsync();   // (straight-sided elements read their constant Jacobians straight from the vertices: no P1a, no barrier)
    HFX_PROF(5);
    // the next element's gather flies while this element is computed
    if (!refCand) prefetchA(e + eStride);

    // ---- P1b: measures, inverses, normals, coefficient interpolation (Operator.cpp:41-84, HDGModel.cpp:53-85, HDGBase.cpp:18-65) --
    for (int k = tid; k < nJ; k += NT) {
      if (k < nIP) {
        const int ip = k;
        double J[DIM][DIM];
#pragma unroll
        for (int r = 0; r < DIM; r++)
#pragma unroll
          for (int m = 0; m < DIM; m++) J[r][m] = aff ? 0.5 * (X[(r + 1) * DIM + m] - X[m]) : JR[ip * D2 + r * DIM + m];
        double det, I[DIM][DIM];
        det_inv(J, det, I);
        {   // is det J constant over the element?  (then M = det J * reference mass exactly, whatever the geometry does otherwise)
          double J0[DIM][DIM];
#pragma unroll
          for (int r = 0; r < DIM; r++)
#pragma unroll
            for (int m = 0; m < DIM; m++) J0[r][m] = aff ? J[r][m] : JR[r * DIM + m];
          const double det0 = aff ? det : det_only(J0);
          if (!(fabs(det - det0) <= 1e-13 * fabs(det0))) QCTR[3] = 1;
          if (ip == 0) {
            FU[ev(nN)] = fast_rcp(det0);
#pragma unroll
            for (int m = 0; m < DIM; m++)
#pragma unroll
              for (int r = 0; r < DIM; r++) GEO[m * DIM + r] = I[m][r];
            GEO[D2] = det;
          }
        }
        const double dv = WQ[ip] * det;
        DV[ip] = dv;
#pragma unroll
        for (int m = 0; m < DIM; m++)
#pragma unroll
          for (int r = 0; r < DIM; r++) IJ[ip * D2 + m * DIM + r] = I[m][r] * dv;
        if (diffField) {
          double Dc[D2];
#pragma unroll
          for (int c = 0; c < D2; c++) Dc[c] = 0.0;
          for (int i = 0; i < nN; i++) {
            const double s = PHI[ip * nNp + i];
#pragma unroll
            for (int c = 0; c < D2; c++) Dc[c] = fma(s, DN[i * D2 + c], Dc[c]);
          }
#pragma unroll
          for (int c = 0; c < D2; c++) DIP[ip * D2 + c] = Dc[c];
        }
        if (hasConv) {
          double v[DIM];
#pragma unroll
          for (int d = 0; d < DIM; d++) v[d] = 0.0;
          for (int i = 0; i < nN; i++) {
            const double s = PHI[ip * nNp + i];
#pragma unroll
            for (int d = 0; d < DIM; d++) v[d] = fma(s, VN[i * DIM + d], v[d]);
          }
#pragma unroll
          for (int d = 0; d < DIM; d++) VIP[ip * DIM + d] = v[d];
        }
        double lw = 0.0;
        if (hasReac) lw += ts * p.reacIP[(size_t)e * nIP + ip] * dv;
        if (euler) lw += dv;
        LW[ip] = lw;
        double rw = hasSrc ? ts * p.srcIP[(size_t)e * nIP + ip] * dv : 0.0;
        if (euler) {   // Mass * Solution_old = sum_ip dV phi_i(ip) u_old(ip)
          const double* so = p.solOld + (size_t)e * nN;
          double uo = 0.0;
          for (int i = 0; i < nN; i++) uo = fma(PHI[ip * nNp + i], so[i], uo);
          rw = fma(dv, uo, rw);
        }
        LW[nIP + ip] = rw;
      } else {
        const int fi = k - nIP, f = fi / nIPf, ip = fi % nIPf;
        const int* fn = FN + f * t;
        double J[DIM - 1][DIM];
#pragma unroll
        for (int r = 0; r < DIM - 1; r++)
#pragma unroll
          for (int m = 0; m < DIM; m++) J[r][m] = aff ? 0.5 * (X[fn[r + 1] * DIM + m] - X[fn[0] * DIM + m]) : JR[(nIP + fi) * D2 + r * DIM + m];
        double tauip = 0.0, Dc[D2], v[DIM];
#pragma unroll
        for (int c = 0; c < D2; c++) Dc[c] = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; d++) v[d] = 0.0;
        for (int a = 0; a < t; a++) {
          const int nd = fn[a];
          const double s = FSH[ip * t + a];
          tauip = fma(s, TAU[f * t + a], tauip);
          if (diffField) {
#pragma unroll
            for (int c = 0; c < D2; c++) Dc[c] = fma(s, DN[nd * D2 + c], Dc[c]);
          }
          if (hasConv) {
#pragma unroll
            for (int d2 = 0; d2 < DIM; d2++) v[d2] = fma(s, VN[nd * DIM + d2], v[d2]);
          }
        }
        double nv[DIM], area;
        if (DIM == 2) {
          nv[0] = -J[0][1]; nv[1] = J[0][0];
          area = sqrt(J[0][0] * J[0][0] + J[0][1] * J[0][1]);
        } else {
          nv[0] = J[0][1] * J[DIM - 2][2 % DIM] - J[0][2 % DIM] * J[DIM - 2][1];
          nv[1] = J[0][2 % DIM] * J[DIM - 2][0] - J[0][0] * J[DIM - 2][2 % DIM];
          nv[DIM - 1] = J[0][0] * J[DIM - 2][1] - J[0][1] * J[DIM - 2][0];
          // sqrt(det(J J^T)) (Operator.cpp:66-69)
          const double g00 = J[0][0] * J[0][0] + J[0][1] * J[0][1] + J[0][2 % DIM] * J[0][2 % DIM];
          const double g11 = J[DIM - 2][0] * J[DIM - 2][0] + J[DIM - 2][1] * J[DIM - 2][1] + J[DIM - 2][2 % DIM] * J[DIM - 2][2 % DIM];
          const double g01 = J[0][0] * J[DIM - 2][0] + J[0][1] * J[DIM - 2][1] + J[0][2 % DIM] * J[DIM - 2][2 % DIM];
          area = sqrt(g00 * g11 - g01 * g01);
        }
        // |J_0 x J_1| = sqrt(det(J J^T)) (2D: |tangent|): the norm of the un-normalised normal is the face measure
        // outward orientation: (x_opposite - x_v0) . n <= 0   (HDGBase.cpp:43-62)
        const int v0 = fn[0];
        const int vn = OPP[f];
        const double inrm = fast_rcp(area);
        double prod = 0.0;
#pragma unroll
        for (int m = 0; m < DIM; m++) { nv[m] *= inrm; prod = fma(X[vn * DIM + m] - X[v0 * DIM + m], nv[m], prod); }
        if (prod > 0.0) {
#pragma unroll
          for (int m = 0; m < DIM; m++) nv[m] = -nv[m];
        }
        if (ip == 0) {
#pragma unroll
          for (int m = 0; m < DIM; m++) GEO[D2 + 1 + f * (DIM + 1) + m] = nv[m];
          GEO[D2 + 1 + f * (DIM + 1) + DIM] = area;
        }
        const double dvf = WQ[nIP + ip] * area;
        double* wt = FWT + (size_t)ip * ldw + f;          // column (kind, f), kind-major: whole column tiles of unused kinds are skipped
        wt[kTau * nFc] = dvf * tauip;
        double vdn = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; d++) {
          wt[(kN + d) * nFc] = -dvf * nv[d];            // the n_d and (Dn)_d mass matrices are stored NEGATED: that is how Sql, Slq, Suq use them
          double dn = nv[d];
          if (diffField) {
            dn = 0.0;
#pragma unroll
            for (int b = 0; b < DIM; b++) dn = fma(Dc[b * DIM + d], nv[b], dn);  // (D n)_d, D col-major
          }
          wt[(kDN + d) * nFc] = hasDiff ? -dvf * dn : 0.0;
          vdn = fma(v[d], nv[d], vdn);
        }
        wt[kC * nFc] = hasConv ? dvf * vdn : 0.0;
        wt[kOne * nFc] = dvf;
      }
    }
    if (aff) {
      // Straight-sided element, beside the per-point work above (done by the first warps only): the purely geometric bulk blocks as
      // scalar combinations of reference matrices (read from L2; their latency hides behind the serial chains of the geometry):
      //   Squ_d = detJ sum_r Jinv(d,r) S^_r  (-> Suq_d bulk part when D = I),   A_d = Sqq^-1 Squ_d = sum_r Jinv(d,r) A^_r.
      constexpr int TOFF = (((nJ + 31) / 32) * 32 <= NT - 64) ? ((nJ + 31) / 32) * 32 : 0;
      if (tid >= TOFF) {
        double J[DIM][DIM], det, Ii[DIM][DIM];
#pragma unroll
        for (int r = 0; r < DIM; r++)
#pragma unroll
          for (int m = 0; m < DIM; m++) J[r][m] = 0.5 * (X[(r + 1) * DIM + m] - X[m]);
        det_inv(J, det, Ii);
        if (affAB && p.srefT) {
          // staged reference matrices, combined in place: A^_r -> A_d, S^_r^T -> Suq_d bulk part (D = c I)
          constexpr int NA2 = nN * nNp / 2;
          double2* const A2 = reinterpret_cast<double2*>(A);
          double2* const SUQ2 = reinterpret_cast<double2*>(SUQ);
          for (int idx2 = tid - TOFF; idx2 < NA2; idx2 += NT - TOFF) {
            double2 ar[DIM], sr[DIM];
#pragma unroll
            for (int r = 0; r < DIM; r++) { ar[r] = A2[r * NA2 + idx2]; sr[r] = SUQ2[r * NA2 + idx2]; }
#pragma unroll
            for (int d = 0; d < DIM; d++) {
              double2 va = make_double2(0.0, 0.0), vs = make_double2(0.0, 0.0);
#pragma unroll
              for (int r = 0; r < DIM; r++) {
                va.x = fma(Ii[d][r], ar[r].x, va.x); va.y = fma(Ii[d][r], ar[r].y, va.y);
                vs.x = fma(Ii[d][r], sr[r].x, vs.x); vs.y = fma(Ii[d][r], sr[r].y, vs.y);
              }
              A2[d * NA2 + idx2] = va;
              SUQ2[d * NA2 + idx2] = make_double2(dsc * det * vs.x, dsc * det * vs.y);
            }
          }
        } else
        for (int idx = tid - TOFF; idx < nN * nNp; idx += NT - TOFF) {
          double sr[DIM], ar[DIM];
#pragma unroll
          for (int r = 0; r < DIM; r++) { sr[r] = __ldg(p.sref + r * nN * nNp + idx); ar[r] = affAB ? __ldg(p.aref + r * nN * nNp + idx) : 0.0; }
          const int k = idx / nNp, n = idx - k * nNp;
#pragma unroll
          for (int d = 0; d < DIM; d++) {
            double vs = 0.0, va = 0.0;
#pragma unroll
            for (int r = 0; r < DIM; r++) { vs = fma(Ii[d][r], sr[r], vs); va = fma(Ii[d][r], ar[r], va); }
            vs *= det;
            if (!affAB) SQU[(d * nN + k) * nNp + n] = vs;                                 // right operand of A = W Squ (general P4 path)
            if (!diffField && n < nN) SUQ[(d * nN + n) * nNp + k] = dsc * vs;   // Suq_d bulk part with D = c I
            if (affAB) A[d * nN * nNp + idx] = va;                                        // A_d column-major: idx = n * nNp + m
          }
        }
      }
    }
    gsync();
    HFX_PROF(1);

    // ---- P2: g[ip][(d,i)] = dV (J^-1 grad_ref phi_i)_d ; cg = suu left operand ----------------------------------------
    if (!needG) {
      if (needSuu) for (int idx = tid; idx < nIP * nN; idx += NT) { const int ip = idx / nN, i = idx % nN; CG[ip * nNp + i] = LW[ip] * PHI[ip * nNp + i]; }
    } else for (int idx = tid; idx < nIP * nN; idx += NT) {
      const int ip = idx / nN, i = idx % nN;
      const double* d = DSH + (ip * nN + i) * DIM;
      double dr[DIM], gg[DIM];
#pragma unroll
      for (int r = 0; r < DIM; r++) dr[r] = d[r];
#pragma unroll
      for (int m = 0; m < DIM; m++) {
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < DIM; r++) s = fma(IJ[ip * D2 + m * DIM + r], dr[r], s);
        gg[m] = s;
        if (!aff || diffField) G[ip * ldg + m * nN + i] = s;   // straight-sided: A (which aliases g) is already in place
      }
      double c = LW[ip] * PHI[ip * nNp + i];
      if (hasConv) {
#pragma unroll
        for (int m = 0; m < DIM; m++) c = fma(-ts * VIP[ip * DIM + m], gg[m], c);
      }
      CG[ip * nNp + i] = c;
    }
    gsync();
    HFX_PROF(2);

    // =====================================================================================================================
    // From here on every product runs on the FP64 tensor cores: a phase is a list of warp tasks (8 output rows x up to 3 column
    // tiles of 8), operands are fetched from shared memory one 64-bit word per lane.

    // ---- P3a: M = sum_ip dV phi phi^T (Mass.cpp:5-38 / HDGBase.cpp:152) ------------------------------------------------------
    // If det J is constant over the element (every straight-sided simplex), M = det J * M_ref exactly and W = M_ref^-1 / det J:
    // no contraction and no inversion.  Curved elements take the general path.
    const bool constDet = (QCTR[3] == 0);
    if (constDet && !affAB) {
      const double rdet = FU[ev(nN)];
      if (kMHReg) {
#pragma unroll
        for (int q = 0; q < NMH; q++) if (tid + q * NT < nNp * nNp) W[tid + q * NT] = mh[q] * rdet;
      } else {
        for (int i = tid; i < nNp * nNp; i += NT) W[i] = p.mhinv[i] * rdet;
      }
    }
    if (!constDet) for (int task = warp; task < MTN * ((MTN + 2) / 3); task += NWARP) {
      const int mt = task % MTN, ng = task / MTN;
      mma_task<3, KS_IP>(mt, ng * 3, lane,
          [&](int m, int k) { return (k < nIP && m < nN) ? DV[k] * PHI[k * nNp + m] : 0.0; },
          [&](int k, int n) { return (k < nIP && n < nN) ? PHI[k * nNp + n] : 0.0; },
          [&](int m, int n, double v0, double v1) {
            if (m < nN) { if (n < nN) Mm[m + nNp * n] = v0; if (n + 1 < nN) Mm[m + nNp * (n + 1)] = v1; }
          });
    }
    if (!constDet && (nN & 1) && tid == NT - 1) {   // odd size: unit pad diagonal for the 2x2-block Gauss-Jordan (pad row/column are zero)
      for (int j = 0; j < nN; j++) { Mm[nN + nNp * j] = 0.0; Mm[j + nNp * nN] = 0.0; }
      Mm[nN + nNp * nN] = 1.0;
    }
    gsync();
    HFX_PROF(3);

    // ---- P3b: (general path only: W = M^-1 by warps 0-3, block Gauss-Jordan) ; contractions Squ, Suu, face matrices, Fu as
    //            warp tasks: dynamic queue when the inversion runs beside them, static round-robin otherwise ------------------
    if (TPE == kAsmThreads) { if (!constDet && tid < kGJThreads) group_invert<nNp, nNp, kGJThreads>(Mm, Wb, tid, p.status, 1); }
    else if (!constDet) group_invert<nNp, nNp, (TPE < kAsmThreads ? TPE : 32)>(Mm, Wb, tid, p.status, 1 + grp);
    HFX_PROF(14);
    {
      const int lr = lane >> 2, lc = lane & 3;
      constexpr int MROWS = DIM * nN, SQ_MT = (MROWS + 7) / 8, NG_N = (MTN + 2) / 3;   // column groups of 3 tiles over nN
      constexpr int T_SQU = SQ_MT * NG_N, T_SUU = MTN * NG_N;
      constexpr int MR = t * t, FW_MT = (MR + 7) / 8, NC = nFc * NW, FW_NG = ((NC + 7) / 8 + 2) / 3, T_FW = FW_MT * FW_NG;
      constexpr int T_ALL = T_SQU + T_SUU + T_FW + 1;
      int task = constDet ? warp + (aff ? T_SQU : 0) : grab1(&QCTR[0], lane);   // straight-sided: Squ comes from the reference matrices below
      while (task < T_ALL) {
        if (task < T_SQU + T_SUU) {
          // Squ_d[k][j] = sum_ip g[ip][(d,k)] phi[ip][j]  (HDGBase.cpp:150; with D = I also Suq_d, HDGDiffusion.cpp:130-144)
          // Suu (bulk part) = sum_ip cg[ip][i] phi[ip][j]: -C^T (Convection.cpp:5-49) + reaction mass + Euler mass
          const bool isSqu = task < T_SQU;
          const int tk = isSqu ? task : task - T_SQU;
          const int mt = isSqu ? tk % SQ_MT : tk % MTN, ng = isSqu ? tk / SQ_MT : tk / MTN;
          const int m = mt * 8 + lr, mrows = isSqu ? MROWS : nN;
          const double* pa = (isSqu ? G : CG) + imin(m, mrows - 1);
          const double* pb[3];
#pragma unroll
          for (int j = 0; j < 3; j++) pb[j] = PHI + imin((ng * 3 + j) * 8 + lr, nN - 1);
          double c[3][2];
          zero_c(c);
          if (isSqu || needSuu) mma_affine<3, nIP>(c, pa, isSqu ? ldg : nNp, pb, nNp, lc);   // Laplace-type models: bulk Suu = 0
          if (m < mrows) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
#pragma unroll
              for (int h = 0; h < 2; h++) {
                const int n = (ng * 3 + j) * 8 + 2 * lc + h;
                if (n < nN) {
                  if (isSqu) {
                    const int d = m / nN, kk = m % nN;
                    SQU[(d * nN + kk) * nNp + n] = c[j][h];                               // right operand of A = W Squ
                    if (!diffField) SUQ[(d * nN + n) * nNp + kk] = dsc * c[j][h];   // left operand of K, R (D = c I)
                  } else {
                    SUU[m + nNp * n] = c[j][h];
                  }
                }
              }
            }
          }
        } else if (task < T_SQU + T_SUU + T_FW) {
          // weighted face mass matrices FW[(f,kind)][a + tp b] = sum_ip wt[ip][(f,kind)] phi_a phi_b
          const int tk = task - (T_SQU + T_SUU);
          const int mt = tk % FW_MT, ng = tk / FW_MT;
          const int m = mt * 8 + lr;
          const unsigned need3 = (tileNeed >> (ng * 3)) & 7u;   // column tiles of kinds no operator of this model reads are skipped
          if (need3) {
            const double* pa = FFS + imin(m, MR - 1);
            const double* pb[3];
#pragma unroll
            for (int j = 0; j < 3; j++) pb[j] = FWT + imin((ng * 3 + j) * 8 + lr, NC - 1);
            double c[3][2];
            zero_c(c);
            constexpr int KSF = (nIPf + 3) / 4;
#pragma unroll
            for (int ks = 0; ks < KSF; ks++) {
              const int k = ks * 4 + lc, kk = k < nIPf ? k : nIPf - 1;
              const double a = k < nIPf ? pa[kk * MR] : 0.0;
#pragma unroll
              for (int j = 0; j < 3; j++) if (need3 & (1u << j)) dmma(c[j], a, pb[j][kk * ldw]);
            }
            if (m < MR) {
              double* dstm = FW + (m % t) + tp * (m / t);
#pragma unroll
              for (int j = 0; j < 3; j++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                  const int n = (ng * 3 + j) * 8 + 2 * lc + h;
                  const bool nKind = aff && (n / nFc) >= kN && (n / nFc) < kN + DIM;   // written from the reference face mass instead
                  if (n < NC && (need3 & (1u << j)) && !nKind) dstm[((n % nFc) * NW + n / nFc) * FWS] = c[j][h];
                }
              }
            }
          }
        } else {
          // Fu = source (Source.cpp:24-48) + Euler mass * Solution_old (Euler.cpp:29-30), both as sum_ip phi_i(ip) * weight(ip)
          for (int i = lane; i < nN; i += 32) {
            double s2 = 0.0;
            if (hasSrc || euler) for (int ip = 0; ip < nIP; ip++) s2 = fma(PHI[ip * nNp + i], LW[nIP + ip], s2);
            FU[i] = s2;
          }
        }
        task = constDet ? task + NWARP : grab1(&QCTR[0], lane);
      }
    }
    if (aff) {
      // straight-sided element: n_d face matrices = -area n_d M^f (stored negated like the contraction does)
      for (int idx = tid; idx < nFc * FWS; idx += NT) {
        const int f = idx / FWS, ab = idx - f * FWS;
        const double* gf = GEO + D2 + 1 + f * (DIM + 1);
        const double sc = -gf[DIM] * ((affAB && p.srefT) ? Mm[ab] : __ldg(p.mfref + ab));
#pragma unroll
        for (int d = 0; d < DIM; d++) FW[(f * NW + kN + d) * FWS + ab] = sc * gf[d];
      }
    }
    gsync();
    HFX_PROF(4);
    prefetchB(e + eStride);

    // ---- P3c: Suq bulk part with a diffusion field (HDGDiffusion.cpp:31-72,130-144) ----------------------------------------
    if (diffField) {
      for (int idx = tid; idx < nIP * nN; idx += NT) {   // g <- D g in place
        const int ip = idx / nN, i = idx % nN;
        double gg[DIM], o[DIM];
#pragma unroll
        for (int m = 0; m < DIM; m++) gg[m] = G[ip * ldg + m * nN + i];
#pragma unroll
        for (int a = 0; a < DIM; a++) {
          double s2 = 0.0;
#pragma unroll
          for (int b = 0; b < DIM; b++) s2 = fma(DIP[ip * D2 + b * DIM + a], gg[b], s2);
          o[a] = s2;
        }
#pragma unroll
        for (int m = 0; m < DIM; m++) G[ip * ldg + m * nN + i] = o[m];
      }
      gsync();
      constexpr int MROWS = DIM * nN, SQ_MT = (MROWS + 7) / 8, NG_N = (MTN + 2) / 3;
      for (int task = warp; task < SQ_MT * NG_N; task += NWARP) {
        mma_task<3, KS_IP>(task % SQ_MT, (task / SQ_MT) * 3, lane,
            [&](int m, int k) { return (k < nIP && m < MROWS) ? G[k * ldg + m] : 0.0; },
            [&](int k, int n) { return (k < nIP && n < nN) ? PHI[k * nNp + n] : 0.0; },
            [&](int m, int n, double v0, double v1) {
              if (m < MROWS && n < nN) {
                const int d = m / nN, kk = m % nN;
                SUQ[(d * nN + n) * nNp + kk] = v0;
                if (n + 1 < nN) SUQ[(d * nN + n + 1) * nNp + kk] = v1;
              }
            });
      }
      gsync();
    }
    HFX_PROF(6);
    if (affAB) {   // B_d = -(area n_d / detJ) B^_f  (after the face contraction: the staged phi_a phi_b table lives in the B region)
      const double rdet = FU[ev(nN)];
      constexpr int NIB = (nFc * nN * t + NT - 1) / NT;
      double bv[NIB];
#pragma unroll
      for (int it = 0; it < NIB; it++) { const int idx = tid + it * NT; bv[it] = idx < nFc * nN * t ? (p.srefT ? brefS[idx] : __ldg(p.bref + idx)) : 0.0; }
#pragma unroll
      for (int it = 0; it < NIB; it++) {
        const int idx = tid + it * NT;
        if (idx < nFc * nN * t) {
          const int f = idx / (nN * t), rem = idx - f * nN * t, m = rem / t, b = rem - m * t;
          const double* gf = GEO + D2 + 1 + f * (DIM + 1);
          const double sc = -gf[DIM] * rdet * bv[it];
#pragma unroll
          for (int d = 0; d < DIM; d++) B[(d * nN + m) * ldc + f * t + b] = sc * gf[d];
        }
      }
      for (int idx = tid; idx < DIM * nN; idx += NT) { B[idx * ldc + l] = 0.0; B[idx * ldc + l + 1] = 0.0; }  // Q0 column
    }
    // face parts of Suu (+tau mass, HDGBase.cpp:128) and Suq (-(Dn) mass, HDGDiffusion.cpp:121), gather form
    for (int idx = tid; idx < nN * nN; idx += NT) {
      const int i = idx % nN, j = idx / nN;
      double suu = SUU[i + nNp * j], suq[DIM];
#pragma unroll
      for (int d = 0; d < DIM; d++) suq[d] = SUQ[(d * nN + j) * nNp + i];
      for (int f = 0; f < nFc; f++) {
        const int a = NIF[f * nN + i], b = NIF[f * nN + j];
        if (a >= 0 && b >= 0) {
          const double* fw = FW + f * NW * FWS + a + tp * b;
          suu += ts * fw[kTau * FWS];
#pragma unroll
          for (int d = 0; d < DIM; d++) suq[d] += dsc * fw[(kDNe + d) * FWS];
        }
      }
      SUU[i + nNp * j] = suu;
#pragma unroll
      for (int d = 0; d < DIM; d++) SUQ[(d * nN + j) * nNp + i] = ts * suq[d];
    }
    if ((nN & 1) && tid == NT - 1) {   // odd size: unit pad diagonal of K for the 2x2-block Gauss-Jordan
      for (int j = 0; j < nN; j++) { SUU[nN + nNp * j] = 0.0; SUU[j + nNp * nN] = 0.0; }
      SUU[nN + nNp * nN] = 1.0;
    }
    gsync();
    HFX_PROF(7);

    // ---- P4: A_d = W Squ_d (col-major out) ;  B_d = W Sql_d with Sql[(fn_f(a),d),(f,b)] = -N_fd[a][b] (HDGBase.cpp:134) ------
    {
      const int lr = lane >> 2, lc = lane & 3;
      constexpr int NG_N = (MTN + 2) / 3, T_A = DIM * MTN * NG_N;
      constexpr int NB = DIM * t, NBT = (NB + 7) / 8, NG_B = (NBT + 2) / 3, T_B = nFc * MTN * NG_B;
      if (!affAB) for (int task = warp; task < T_A + T_B; task += NWARP) {
        if (task < T_A) {
          const int d = task / (MTN * NG_N), r = task % (MTN * NG_N);
          const int m = (r % MTN) * 8 + lr, ng = r / MTN;
          const double* pa = W + imin(m, nN - 1);
          const double* pb[3];
#pragma unroll
          for (int j = 0; j < 3; j++) pb[j] = SQU + d * nN * nNp + imin((ng * 3 + j) * 8 + lr, nN - 1);
          double c[3][2];
          zero_c(c);
          mma_affine<3, nN>(c, pa, nNp, pb, nNp, lc);
          if (m < nN) {
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
              for (int h = 0; h < 2; h++) {
                const int n = (ng * 3 + j) * 8 + 2 * lc + h;
                if (n < nN) A[(d * nN + n) * nNp + m] = c[j][h];
              }
          }
        } else {
          const int tb = task - T_A, f = tb / (MTN * NG_B), r = tb % (MTN * NG_B);
          const int m = (r % MTN) * 8 + lr, ng = r / MTN;
          const int* fn = FN + f * t;
          // left operand: gathered columns W[:, fn_f(k)] ; right operand: -N_fd[k][b] = FW[(f,kN+d)][k + tp b], column n = (d, b)
          const double* wrow = W + imin(m, nN - 1);
          const double* pb[3];
#pragma unroll
          for (int j = 0; j < 3; j++) {
            const int n = imin((ng * 3 + j) * 8 + lr, NB - 1);
            pb[j] = FW + (f * NW + kN + n / t) * FWS + tp * (n % t);
          }
          double c[3][2];
          zero_c(c);
          constexpr int KS = (t + 3) / 4;
#pragma unroll
          for (int ks = 0; ks < KS; ks++) {
            const int k = ks * 4 + lc, kk = k < t ? k : t - 1;
            const double a = k < t ? wrow[nNp * fn[kk]] : 0.0;
#pragma unroll
            for (int j = 0; j < 3; j++) if ((ng * 3 + j) < NBT) dmma(c[j], a, pb[j][kk]);
          }
          if (m < nN) {
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
              for (int h = 0; h < 2; h++) {
                const int n = (ng * 3 + j) * 8 + 2 * lc + h;
                if (n < NB) B[NLUT[n] + m * ldc + f * t] = c[j][h];
              }
          }
        }
      }
      if (!affAB) for (int idx = tid; idx < DIM * nN; idx += NT) { B[idx * ldc + l] = 0.0; B[idx * ldc + l + 1] = 0.0; }  // Q0 column
    }
    if (!affAB) gsync();
    HFX_PROF(8);
    }   // !ref

    // ---- P5: K = Suu - sum_d Suq_d A_d (HDGSolver.cpp:335), R = Sul - sum_d Suq_d B_d with column l = -Fu (:342-343).
    //      A warp task = one column tile x every row tile: the right operand is fetched once per reduction step and feeds MTN
    //      independent accumulator chains (operand traffic, not the DMMA pipe, is what limits these skinny products).
    double* const KB = (W == Mm) ? Wb : Mm;          // the buffer that does not hold W (W is dead once A and B exist)
    constexpr bool kBlock4 = (TPE == kAsmThreads) && (nNp % 4 == 0) && (nNp >= 16);   // 4x4-block pivots with tensor-core updates (block4_invert)
    const bool useB4 = kBlock4 && p.gjThreads != 512;
    double* const KI = useB4 ? (((nNp / 4) & 1) ? KB : SUU) : (((nNp / 2) & 1) ? KB : SUU);
    double* const KC = W;                            // copy of K for the refinement step of U (P7r); the Gauss-Jordan consumes SUU and KB
    {
      const int lr = lane >> 2, lc = lane & 3;
      constexpr int LT = (l + 7) / 8, T_K = MTN, T_R = LT;
      for (int task = warp; task < T_K + T_R; task += NWARP) {
        const bool isK = task < T_K;
        const int nt = isK ? task : task - T_K;
        const int ncl = imin(nt * 8 + lr, isK ? nN - 1 : l);
        const double* pb = isK ? A + ncl * nNp : B + ncl;                          // K: A_d[k'][n] column-major ; R: B_d[k'][n] row-major
        const int sb = isK ? 1 : ldc, sd = isK ? nN * nNp : nN * ldc;
        const double* pa[MTN];
#pragma unroll
        for (int i = 0; i < MTN; i++) pa[i] = SUQ + imin(i * 8 + lr, nN - 1);     // Suq[m][(d,k')]: stride nNp in k'
        double c[MTN][2];
        zero_c(c);
#pragma unroll
        for (int d = 0; d < DIM; d++) {
#pragma unroll
          for (int ks = 0; ks < KS_N; ks++) {
            const int k = ks * 4 + lc, kk = (ks * 4 + 3 < nN) ? k : imin(k, nN - 1);
            const bool dead = (ks * 4 + 3 >= nN) && k >= nN;
            const double b = pb[d * sd + kk * sb];
#pragma unroll
            for (int i = 0; i < MTN; i++) { const double a = pa[i][(d * nN + kk) * nNp]; dmma(c[i], dead ? 0.0 : a, b); }
          }
        }
#pragma unroll
        for (int i = 0; i < MTN; i++) {
          const int m = i * 8 + lr;
          if (m < nN) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const int cc = nt * 8 + 2 * lc + h;
              const double v = c[i][h];
              if (isK) { if (cc < nN) { const double kv = SUU[m + nNp * cc] - v; SUU[m + nNp * cc] = kv; KC[m + nNp * cc] = kv; } }
              else if (cc < l) {
                const int f = cc / t, b = cc % t, a = NIF[f * nN + m];
                double sul = 0.0;
                if (a >= 0) { const double* fw = FW + f * NW * FWS + a + tp * b; sul = ts * ((hasConv ? fw[kC * FWS] : 0.0) - fw[kTau * FWS]); }
                R[m * ldc + cc] = sul - v;
              }
            }
          }
        }
      }
      for (int i = tid; i < nN; i += NT) { R[i * ldc + l] = -FU[i]; R[i * ldc + l + 1] = 0.0; }
    }
    gsync();
    HFX_PROF(9);
    // ---- P6: K^-1 (2x2-block-pivot Gauss-Jordan on all eight warps: the pivot chain is serial, measured variants that ran it on
    //      four warps or on one warp beside the R tiles were slower) ------------------------------------------------
    if (kBlock4 && useB4) { block4_invert<kBlock4 ? nNp : 4, nNp>(SUU, KB, tid, p.status, 1 + grp); gsync(); }
    else group_invert<nNp, nNp, NT>(SUU, KB, tid, p.status, 1 + grp);
    HFX_PROF(10);

    // ---- P7: U = -K^-1 R ; U0 = K^-1 Fu (column l), then one step of iterative refinement, U <- U - K^-1 (K U + R).
    //      U = -K^-1 R through an EXPLICIT inverse carries a forward error of eps cond(K) ||K^-1|| ||R||, and ||K^-1|| ||R|| / ||U|| ~ 1 / h on
    //      fine meshes (K^-1 is dominated by the constant mode, which R barely excites); the refined U has the eps cond(K) ||U|| of the
    //      reference's triangular solves (HDGSolver.cpp:331-343, HouseholderQR).  A warp owns one column tile through all three products
    //      (a column of U depends on the same column of R only), so the steps are separated by warp barriers, not block barriers.
    {
      const int lr = lane >> 2, lc = lane & 3;
      constexpr int L1T = (l + 1 + 7) / 8;
      // Cm[:, tile] = beta * Cm[:, tile] + sgn * Am * Bm[:, tile]
      auto col_task = [&](int task, const double* Am, const double* Bm, double* Cm, double sgn, double beta) {
        const double* pb = Bm + imin(task * 8 + lr, l);
        const double* pa[MTN];
#pragma unroll
        for (int i = 0; i < MTN; i++) pa[i] = Am + imin(i * 8 + lr, nN - 1);
        double c[MTN][2];
        zero_c(c);
#pragma unroll
        for (int ks = 0; ks < KS_N; ks++) {
          const int k = ks * 4 + lc, kk = (ks * 4 + 3 < nN) ? k : imin(k, nN - 1);
          const bool dead = (ks * 4 + 3 >= nN) && k >= nN;
          const double b = pb[kk * ldc];
#pragma unroll
          for (int i = 0; i < MTN; i++) { const double a = pa[i][kk * nNp]; dmma(c[i], dead ? 0.0 : a, b); }
        }
#pragma unroll
        for (int i = 0; i < MTN; i++) {
          const int m = i * 8 + lr, n = task * 8 + 2 * lc;
          if (m < nN && n <= l) {
            double2* d2 = reinterpret_cast<double2*>(Cm + m * ldc + n);
            const double2 o = beta != 0.0 ? *d2 : make_double2(0.0, 0.0);
            *d2 = make_double2(fma(sgn, c[i][0], o.x), fma(sgn, c[i][1], o.y));
          }
        }
      };
      for (int task = warp; task < L1T; task += NWARP) {
        col_task(task, KI, R, Um, -1.0, 0.0);    // U = -K^-1 R
        __syncwarp();
        col_task(task, KC, Um, R, 1.0, 1.0);     // V = R + K U   (in place over the tile's columns of R)
        __syncwarp();
        col_task(task, KI, R, Um, -1.0, 1.0);    // U -= K^-1 V
      }
    }
    gsync();
    HFX_PROF(11);

    // ---- P8: Q_d = -A_d U - B_d ; Q0_d = -A_d U0  (:344-345) -----------------------------------------------------------
    //      warp task = (d, column tile) x every row tile; the tasks of the last, partial round are split into single tiles
    {
      const int lr = lane >> 2, lc = lane & 3;
      constexpr int L1T = (l + 1 + 7) / 8, NT8 = DIM * L1T, NFULL = (NT8 / NWARP) * NWARP, NREM = NT8 - NFULL;
      auto q_task = [&](int tk, int iBeg, int iEnd) {
        const int d = tk / L1T, nt = tk - d * L1T;
        const double* pb = Um + imin(nt * 8 + lr, l);
        const double* pa[MTN];
#pragma unroll
        for (int i = 0; i < MTN; i++) pa[i] = A + d * nN * nNp + imin(i * 8 + lr, nN - 1);     // A_d[m][k], column-major
        double c[MTN][2];
        zero_c(c);
#pragma unroll
        for (int ks = 0; ks < KS_N; ks++) {
          const int k = ks * 4 + lc, kk = (ks * 4 + 3 < nN) ? k : imin(k, nN - 1);
          const bool dead = (ks * 4 + 3 >= nN) && k >= nN;
          const double b = pb[kk * ldc];
#pragma unroll
          for (int i = 0; i < MTN; i++) if (i >= iBeg && i < iEnd) { const double a = pa[i][kk * nNp]; dmma(c[i], dead ? 0.0 : a, b); }
        }
#pragma unroll
        for (int i = 0; i < MTN; i++) {
          const int m = i * 8 + lr, n = nt * 8 + 2 * lc;
          if (i >= iBeg && i < iEnd && m < nN && n <= l) {
            double2* bq = reinterpret_cast<double2*>(B + (d * nN + m) * ldc + n);
            const double2 o = *bq;
            *bq = make_double2(-c[i][0] - o.x, -c[i][1] - o.y);
          }
        }
      };
      for (int task = warp; task < NFULL; task += NWARP) q_task(task, 0, MTN);
      if (NREM > 0) for (int st = warp; st < NREM * MTN; st += NWARP) q_task(NFULL + st / MTN, st % MTN, st % MTN + 1);
    }
    if (kBulkUQ) fence_proxy_async();
    gsync();
    HFX_PROF(12);
    // U, Q leave as whole rows (row-major per element in HBM): one bulk copy per row, issued here so that they fly during P9
    if (kBulkUQ) {
      constexpr int q = DIM * nN;
      constexpr int RPWQ = (nN + q + NWARP - 1) / NWARP;   // a warp issues its copies one after the other: spread them evenly
#pragma unroll
      for (int r0 = 0; r0 < RPWQ; r0 += 32) {
        const int row = warp * RPWQ + r0 + lane;
        if (r0 + lane < RPWQ && row < nN + q) {
          if (row < nN) bulk_store(p.U + ((size_t)e * nN + row) * l, Um + row * ldc, l * 8);
          else { const int rq = row - nN; bulk_store(p.Q + ((size_t)e * q + rq) * l, B + ((rq % DIM) * nN + rq / DIM) * ldc, l * 8); }
          bulk_commit();
        }
      }
    }

    // ---- P9z (all-reference path): every face operator of the element is a multiple of the face mass (Slu = tau_f area_f M^f,
    //      Slq_d = -c area_f n_fd M^f, Sll = -Slu), so the rows are combined first,
    //          Z_f[b][:] = tau_f (U[fn_f(b)][:] - I) - c sum_d n_fd Q_d[fn_f(b)][:],
    //      and S_f = (area_f M^f) Z_f is a product with reduction length t instead of (1 + DIM) t: 144 DMMAs instead of 480 at p=3.
    //      Z_f lives in the (D n)_d / v.n slots of the face's weight matrices, which this path does not use; its S0 column in Suq.
    const bool zPath = ref && kBulkS;
    double* const Z0 = SUQ;
    if (zPath) {
      for (int item = tid; item < l * (l / 2); item += NT) {
        const int fb = item / (l / 2), c = 2 * (item - fb * (l / 2)), f = fb / t, b = fb - f * t;
        const int nd = FN[fb];
        const double* gf = GEO + D2 + 1 + f * (DIM + 1);
        const double tf = TAU[f * t];
        const double2 uu = *reinterpret_cast<const double2*>(Um + nd * ldc + c);
        double zx = tf * (uu.x - (c == fb ? 1.0 : 0.0)), zy = tf * (uu.y - (c + 1 == fb ? 1.0 : 0.0));
#pragma unroll
        for (int d = 0; d < DIM; d++) {
          const double2 qq = *reinterpret_cast<const double2*>(B + (d * nN + nd) * ldc + c);
          const double w = -dsc * gf[d];
          zx = fma(w, qq.x, zx); zy = fma(w, qq.y, zy);
        }
        *reinterpret_cast<double2*>(FW + (f * NW + kDN) * FWS + b * l + c) = make_double2(zx, zy);
      }
      for (int fb = tid; fb < l; fb += NT) {
        const int f = fb / t, nd = FN[fb];
        const double* gf = GEO + D2 + 1 + f * (DIM + 1);
        double z = TAU[f * t] * Um[nd * ldc + l];
#pragma unroll
        for (int d = 0; d < DIM; d++) z = fma(-dsc * gf[d], B[(d * nN + nd) * ldc + l], z);
        Z0[fb] = z;
      }
      gsync();
    }
    // ---- P9: S = Slu U + Slq Q + Sll ; S0 = Fl - Slu U0 - Slq Q0 (:347-348); Dirichlet rows (:489-501); scatter (:596-618) --
    //      a warp task = all row tiles of one face x three column tiles: one gathered right-operand load feeds every row tile
    {
      const int lr = lane >> 2, lc = lane & 3;
      constexpr int TT = (t + 7) / 8, L1T = (l + 1 + 7) / 8, NTW9 = 3, NG = (L1T + NTW9 - 1) / NTW9;
      constexpr int KTOT = (1 + DIM) * t, KS_S = (KTOT + 3) / 4;
      for (int task = warp; task < nFc * NG; task += NWARP) {
        const int f = task / NG, ng = task - f * NG;
        const double* fwf = FW + f * NW * FWS;
        const unsigned short* klut = KLUT + 2 * f * L::nKLUT;
        // bulk write-out: output rows / columns are face-node POSITIONS; the operands are fetched at the element-local indices they map to
        int acl[TT], ncl[NTW9];
#pragma unroll
        for (int i = 0; i < TT; i++) { acl[i] = imin(i * 8 + lr, t - 1); if (kBulkS) acl[i] = CMAP[f * t + acl[i]] - f * t; }
#pragma unroll
        for (int j = 0; j < NTW9; j++) { ncl[j] = imin((ng * NTW9 + j) * 8 + lr, l); if (kBulkS) ncl[j] = CMAP[ncl[j]]; }
        double c[TT][NTW9][2];
#pragma unroll
        for (int i = 0; i < TT; i++) zero_c(c[i]);
        if (zPath) {
          constexpr int KS_Z = (t + 3) / 4;
          const double* zb[NTW9]; int zs[NTW9];
#pragma unroll
          for (int j = 0; j < NTW9; j++) { const bool in = ncl[j] < l; zb[j] = in ? FW + (f * NW + kDN) * FWS + ncl[j] : Z0 + f * t; zs[j] = in ? l : 1; }
#pragma unroll
          for (int ks = 0; ks < KS_Z; ks++) {
            const int k = ks * 4 + lc, kk = imin(k, t - 1);
            double av[TT], bv[NTW9];
#pragma unroll
            for (int i = 0; i < TT; i++) av[i] = k < t ? fwf[kOne * FWS + tp * kk + acl[i]] : 0.0;
#pragma unroll
            for (int j = 0; j < NTW9; j++) bv[j] = zb[j][kk * zs[j]];
#pragma unroll
            for (int i = 0; i < TT; i++)
#pragma unroll
              for (int j = 0; j < NTW9; j++) dmma(c[i][j], av[i], bv[j]);
          }
        } else
#pragma unroll
        for (int ks = 0; ks < KS_S; ks++) {
          const int k = ks * 4 + lc;
          const int2 off = make_int2((int)klut[2 * k], (int)klut[2 * k + 1]);
          const bool live = (k < KTOT) && (k < t || hasDiff);
          double av[TT], bv[NTW9];
          const double asc = k < t ? 1.0 : dsc;   // Slq = -(D n) mass: the n_d masses scaled by c when D = c I
#pragma unroll
          for (int i = 0; i < TT; i++) av[i] = live ? asc * fwf[off.x + acl[i]] : 0.0;
          const double* brow = sm + off.y;
#pragma unroll
          for (int j = 0; j < NTW9; j++) bv[j] = brow[ncl[j]];
          if (kBulkS && ks * 4 < t) {   // Sll = -tau mass (HDGBase.cpp:125) rides on the tau-mass steps: tau mass (U - I)
#pragma unroll
            for (int j = 0; j < NTW9; j++) bv[j] -= (k < t && ncl[j] == f * t + k) ? 1.0 : 0.0;
          }
#pragma unroll
          for (int i = 0; i < TT; i++)
#pragma unroll
            for (int j = 0; j < NTW9; j++) dmma(c[i][j], av[i], bv[j]);
        }
        const int bcf = BCF[f];
#pragma unroll
        for (int i = 0; i < TT; i++) {
          const int a = i * 8 + lr;
          if (a < t) {
            if (!kBulkS) {
              double* strow = ST + (f * t + a) * ldc;
#pragma unroll
              for (int j = 0; j < NTW9; j++) {
                const int cc = (ng * NTW9 + j) * 8 + 2 * lc;
                if (cc <= l) *reinterpret_cast<double2*>(strow + cc) = make_double2(c[i][j][0], c[i][j][1]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < NTW9; j++) {
                const int cp = (ng * NTW9 + j) * 8 + 2 * lc;   // (a, cp): positions
                if (cp < l) {
                  const int f2 = cp / t, pb = cp - f2 * t;
                  double v0 = c[i][j][0], v1 = c[i][j][1];
                  if (hasConv || bcf) {   // convection part of Sll (HDGConvection.cpp:96), boundary rows (HDGSolver.cpp:489-501)
                    const double* fwd = FW + f * NW * FWS + acl[i];
                    const int b0 = CMAP[cp] - f2 * t, b1 = CMAP[cp + 1] - f2 * t;
                    if (f2 == f) {
                      if (hasConv) { v0 += fwd[kC * FWS + tp * b0]; v1 += fwd[kC * FWS + tp * b1]; }
                      if (bcf == 2) { v0 = fwd[kOne * FWS + tp * b0]; v1 = fwd[kOne * FWS + tp * b1]; }   // IntegratedDirichletModel row
                    } else if (bcf == 2) { v0 = 0.0; v1 = 0.0; }
                    if (bcf == 1) { v0 = (f2 == f && pb == a) ? 1.0 : 0.0; v1 = (f2 == f && pb + 1 == a) ? 1.0 : 0.0; }   // DirichletModel row (Set)
                  }
                  *reinterpret_cast<double2*>(ST + ((f * nFc + f2) * t + a) * t + pb) = make_double2(v0, v1);
                } else if (cp == l) ST[l * l + f * t + a] = c[i][j][0];
              }
            }
          }
        }
      }
    }
    if (kBulkS) fence_proxy_async();
    gsync();
    HFX_PROF(13);

    // ---- P10: write-out.  S: every (face, neighbour face) block of the element is one contiguous t x t block of the global block CSR,
    //      already final and in face-node order in the staging area: one bulk copy each, a bulk reduce-add (f64) where the second element of
    //      an interior face adds to the same diagonal block (two contributors, zeroed storage: the sum is order independent).
    //      U, Q rows left after P8.  Elements whose sizes break the 16-byte granularity take the per-entry path.
    {
      constexpr int q = DIM * nN;
      if (!kBulkUQ) {
        double* gU = p.U + (size_t)e * nN * l;
        for (int idx = tid; idx < nN * l; idx += NT) { const int r = idx / l, c = idx - r * l; gU[idx] = Um[r * ldc + c]; }
        double* gQ = p.Q + (size_t)e * q * l;
        for (int idx = tid; idx < q * l; idx += NT) { const int rq = idx / l, c = idx - rq * l; gQ[idx] = B[((rq % DIM) * nN + rq / DIM) * ldc + c]; }
      }
      double* gS = p.S ? p.S + (size_t)e * l * l : nullptr;
      double* gS0 = p.S0 ? p.S0 + (size_t)e * l : nullptr;
      if (kBulkS) {
        constexpr int OPW = (nFc * nFc + NWARP - 1) / NWARP;
        const int k = warp * OPW + lane;
        if (lane < OPW && k < nFc * nFc) {
          const int f = k / nFc, f2 = k - f * nFc;
          const double* src = ST + k * t * t;
          double* dst = p.vals + ROWS[f] + (long long)POS[k] * t * t;
          if (f2 == f && INTF[f]) bulk_add_f64(dst, src, t * t * 8); else bulk_store(dst, src, t * t * 8);
          bulk_commit();
        }
        if (gS) for (int idx = tid; idx < l * l; idx += NT) {
          const int cc = idx / l, r = idx - cc * l, f = r / t, f2 = cc / t;
          gS[idx] = ST[((f * nFc + f2) * t + PERM[r]) * t + PERM[cc]];
        }
      } else {
        constexpr int RPW = (l + NWARP - 1) / NWARP;
#pragma unroll
        for (int h = 0; h < (l + 31) / 32; h++) {
          const int cc = lane + 32 * h;
          if (cc < l) {
            const int f2 = cc / t, b2 = cc - f2 * t;
            double sv[RPW];
            int off[RPW];
#pragma unroll
            for (int i = 0; i < RPW; i++) {
              const int r = warp + NWARP * i;
              if (r < l) {
                const int f = r / t, a = r - f * t, bc = BCF[f];
                const double* fwd = FW + f * NW * FWS + a + tp * b2;
                const bool diag = (f2 == f);
                double v = ST[r * ldc + cc];
                const double dterm = (hasConv ? fwd[kC * FWS] : 0.0) - fwd[kTau * FWS];   // Sll = -tau mass + (v.n) mass
                v += diag ? dterm : 0.0;
                if (bc == 1) v = (diag && b2 == a) ? 1.0 : 0.0;                            // DirichletModel row (Set)
                else if (bc == 2) v = diag ? fwd[kOne * FWS] : 0.0;                        // IntegratedDirichletModel row
                sv[i] = v;
                off[i] = POSROW[f * l + cc];
              }
            }
#pragma unroll
            for (int i = 0; i < RPW; i++) {
              const int r = warp + NWARP * i;
              if (r < l) {
                const int f = r / t;
                if (gS) gS[r + (size_t)l * cc] = sv[i];
                double* dst = p.vals + RBASE[r] + off[i];
                if (f2 == f && INTF[f]) atomicAdd(dst, sv[i]); else *dst = sv[i];
              }
            }
          }
        }
      }
      if (tid < nN) p.U0[(size_t)e * nN + tid] = Um[tid * ldc + l];
      if (NT > 64) {
        if (tid >= 64 && tid < 64 + q) { const int rq = tid - 64; p.Q0[(size_t)e * q + rq] = B[((rq % DIM) * nN + rq / DIM) * ldc + l]; }
        if (q > NT - 64) for (int rq = NT - 64 + tid; rq < q; rq += NT) p.Q0[(size_t)e * q + rq] = B[((rq % DIM) * nN + rq / DIM) * ldc + l];
      } else for (int rq = tid; rq < q; rq += NT) p.Q0[(size_t)e * q + rq] = B[((rq % DIM) * nN + rq / DIM) * ldc + l];
      if (tid < l) {
        const int r = tid, f = r / t, a = r % t, F = ISM[f], bc = BCF[f];
        const double* fwf = FW + f * NW * FWS;
        double s0 = kBulkS ? -ST[l * l + f * t + PERM[r]] : -ST[r * ldc + l];
        if (bc == 1) s0 = p.dirichlet[(size_t)F * t + a];
        else if (bc == 2) {
          s0 = 0.0;
          for (int b = 0; b < t; b++) s0 = fma(fwf[kOne * FWS + a + tp * b], p.dirichlet[(size_t)F * t + b], s0);
        }
        if (gS0) gS0[r] = s0;
        const int rowDof = F * t + PERM[r];
        if (INTF[f]) atomicAdd(p.rhs + rowDof, s0); else p.rhs[rowDof] = s0;
      }
      if (kBulkUQ || kBulkS) bulk_wait_read();   // the staging areas are rewritten by the next element pass
    }
    gsync();
    HFX_PROF(15);
  }
}

// host-side launch helper
template <int DIM, int P, int TPE>
inline cudaError_t launch_assemble_tpe(const AsmParams& p, int nSM, cudaStream_t st) {
  using L = AsmSmem<DIM, P>;
  constexpr int NGRP = kAsmThreads / TPE;
  constexpr size_t bytes = L::gbytes * NGRP;
  {   // the attribute is per device (a process may hold contexts on several GPUs): set on every launch, it costs microseconds
    cudaError_t e = cudaFuncSetAttribute(hdg_assemble_kernel<DIM, P, TPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
  }
  int perSM = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, hdg_assemble_kernel<DIM, P, TPE>, kAsmThreads, bytes);
  if (perSM < 1) perSM = 1;
  if (const char* ev_ = getenv("HFX_CTAS_PER_SM")) { int v = atoi(ev_); if (v >= 1 && v < perSM) perSM = v; }   // experiments only
  long long grid = (long long)nSM * perSM;
  const long long need = ((long long)(p.eEnd - p.eBegin) + NGRP - 1) / NGRP;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  hdg_assemble_kernel<DIM, P, TPE><<<(int)grid, kAsmThreads, bytes, st>>>(p);
  return cudaGetLastError();
}
// Threads per element by element size: a 4-node element does not feed 256 threads; the linear elements get one warp each.
template <int DIM, int P>
inline cudaError_t launch_assemble_t(const AsmParams& p, int nSM, cudaStream_t st) {
  using C = ElemCfg<DIM, P>;
  constexpr int nN = C::nN;
  constexpr int TPE = nN <= 6 ? 32 : nN <= 10 ? (DIM == 3 ? 128 : 64) : nN <= 15 ? 128 : 256;
  if (getenv("HFX_ONE_GROUP")) return launch_assemble_tpe<DIM, P, 256>(p, nSM, st);   // experiments / A-B comparison
  if (DIM == 3 && P == 2 && getenv("HFX_TPE64")) return launch_assemble_tpe<DIM, P, (DIM == 3 && P == 2) ? 64 : TPE>(p, nSM, st);
  return launch_assemble_tpe<DIM, P, TPE>(p, nSM, st);
}

}  // namespace hfx
