// reagent_b200 -- K2 on the 5th-generation tensor cores: the fused DQN TD-target / loss /
// backward step (same contract as rb200_dqn.cu) with every matrix product issued as
// tcgen05.mma (kind::tf32, 3xTF32 error compensation) and the accumulators in Tensor Memory.
//
// Formulation.  A CTA owns 32 batch rows and computes every layer TRANSPOSED:
//     D_l^T [features x 32 rows]  =  W_l [features x K]  .  H_{l-1}^T [K x 32 rows]
// so the WEIGHTS are the UMMA A operand (M = 128 output features per tile, always a full
// tensor-core tile however small the batch tile is) and the ACTIVATIONS are the B operand
// (N = 32).  4096 rows therefore spread over 128 SMs instead of the 32 an M = 128-row tile
// would use, and the accumulator of a layer is only 32 TMEM columns per 128 features.
//
//   weights      pre-tiled once per update (by the Adam kernel, or dqn_tc_pack_kernel) into one
//                fp32 image per (128-feature tile, 32-k chunk) -- [k/4][row][4 floats].  The TD
//                kernel streams the images through a shared-memory ring with 1-D bulk copies
//                (cp.async.bulk -> mbarrier complete_tx, producer warp; a stage = two chunk
//                images); FOUR LOADER WARPS (one per TMEM lane quadrant; a thread owns one weight
//                row) read "my row" with conflict-free 16-byte loads, split it into TF32 hi / lo
//                in registers and store both into a ring of TENSOR MEMORY columns (tcgen05.st):
//                the weights are the A operand of the MMAs FROM TENSOR MEMORY.  Measured on B200
//                (profiles/r02_summary.md): with A in tensor memory an MMA costs exactly its
//                math (N/2 cycles at M = 128, K = 8), with A in shared memory 39-48 cycles
//                whatever N <= 64; every weight byte crosses shared memory once in and once out
//                as raw fp32.  The step is bound by the latency of this producer -> loader ->
//                MMA chain and of the layer hand-overs, not by any pipe.
//   activations  live in shared memory as hi/lo planes in the same canonical layout (rows =
//                batch rows); the epilogue of layer l (tcgen05.ld -> bias -> activation -> split)
//                writes them straight into the B operand of layer l+1 and, for the online
//                pass on `state`, to the global buffers the weight-gradient kernel reads.
//   backward     dZ_{l-1}^T = W_l^T . dZ_l^T uses pre-transposed weight images; the epilogue
//                multiplies by act'(h_{l-1}) and stores dZ_{l-1}.
//
// Every thread of warps 0-7 owns one output feature (TMEM lane) in the epilogues, which makes
// the bias a per-thread scalar and the operand stores bank-conflict free; warp 8 streams the
// weights, warp 9 issues the MMAs (warp-uniform loops, one elected lane issues), warps 10-13
// move the weights from shared to tensor memory.  Synchronisation is mbarrier-only inside the
// step loop:
//     full[s]   bulk copy landed in smem stage s     sfree[s]  the loaders have read stage s
//     afull[t]  TMEM stage t holds a split chunk     adone[t]  its MMAs retired (tcgen05.commit)
//     dready    a layer's accumulator is complete    opready   next B operand is in smem
//
// Reference semantics: reagent/training/dqn_trainer.py:157-239, dqn_trainer_base.py:33-77,
// 216-241 (see rb200_dqn.cu for the line-by-line map; the loss code is the same).
#include <stdlib.h>
#include <string.h>

#include "rb200_dqn_tc_layout.cuh"
#include "rb200_umma.cuh"

namespace rb200 {

constexpr int kQR = 32;                                   // batch rows per CTA
// A ring stage (shared memory and tensor memory alike) holds kQSub consecutive 32-k chunk images
// of one feature tile (they are contiguous in the image: the k-quad sequence simply continues).
// Every mbarrier wait costs 100-150 cycles even when it passes at once, and each agent of the
// weight stream is a sequential loop with two waits per stage, so the stage is the unit that
// amortises them: 64 k per stage halves the synchronisation cost per MMA.
#ifndef RB200_QSUB
#define RB200_QSUB 2
#endif
constexpr int kQSub = RB200_QSUB;
#ifndef RB200_QSTAGES
#define RB200_QSTAGES (6 / RB200_QSUB)
#endif
constexpr int kQStages = RB200_QSTAGES;                   // shared-memory ring depth (raw fp32 stages)
constexpr int kQStageBytes = kQSub * (kQKC / 4) * kQFullLbo;  // kQSub chunk images
constexpr int kAMaxStages = 7;                            // tensor-memory ring depth (split chunks), upper bound:
                                                          // the plan uses every column the accumulators leave free
constexpr int kASubCols = 2 * kQKC;                       // hi columns then lo columns of a chunk
constexpr int kAStageCols = kQSub * kASubCols;
// B operand (activations): per k quad 64 rows of 16 B -- rows 0-31 hold the hi parts of the 32
// batch rows, rows 32-63 their lo parts -- plus 16 B of padding.  One N = 64 MMA against W_hi
// then yields W_hi.X_hi in accumulator columns 0-31 and W_hi.X_lo in columns 32-63; a second
// N = 32 MMA adds W_lo.X_hi to columns 0-31.  Two MMAs per k step instead of three, and the
// epilogue adds the two column groups.
constexpr int kQLboB = 64 * 16 + 16;
constexpr int kQLoOff = 32 * 4;                           // floats from a hi element to its lo
constexpr int kQEpiThreads = 256;
#ifndef RB200_QGROUPS
#define RB200_QGROUPS 1
#endif
// Loader warps: kQLoaderGroups groups of four (one warp per TMEM lane quadrant).  Chunk i of the
// weight stream belongs to group i % kQLoaderGroups, so consecutive chunks are converted by
// different warps concurrently: one warp needs ~350 cycles per chunk (wait, 8 loads, ~100 ALU
// instructions, two tensor-memory stores and their completion), the MMAs of a chunk 192.
constexpr int kQLoaderGroups = RB200_QGROUPS;
constexpr int kQLoaderWarps = 4 * kQLoaderGroups;
// registers per thread: the register file is 16 K per SM sub-partition and the warps of a CTA
// are dealt round-robin to the four sub-partitions
constexpr int kQRegs = kQLoaderGroups >= 3 ? 80 : (kQLoaderGroups == 2 ? 96 : 128);
constexpr int kQThreads = kQEpiThreads + 64 + 32 * kQLoaderWarps;  // + producer + MMA + loaders
constexpr int kQMaxTiles = 4;                             // accumulators: 4 feature tiles x 64 columns
constexpr int kQAccCols = kQMaxTiles * 64;
constexpr int kQTmemCols = 512;                           // accumulators + the weight ring
static_assert(kQAccCols + 2 * kAStageCols <= kQTmemCols, "tensor memory budget");
static_assert(((10 + kQLoaderWarps + 3) / 4) * 32 * kQRegs <= 16384, "register file budget per sub-partition");
static_assert(kQKC == 32, "loader warps move 32-k chunks");
static_assert(kQStages % kQLoaderGroups == 0, "a shared-memory ring stage belongs to one loader group");
constexpr int kQMaxSteps = 4 * kMaxLayers;
constexpr int kQMaxSmem = 232448;
#ifndef RB200_TC_TIMELINE
#define RB200_TC_TIMELINE 0  // 1: build the clock64 timeline / MMA-skipping hooks (profiling only)
#endif
constexpr bool kTimeline = RB200_TC_TIMELINE != 0;                         // 227 KB opt-in limit of sm_100

enum { kStepHidden = 0, kStepLast = 1, kStepBwd = 2 };

struct QStep {
  uint32_t pack_off;  // byte offset of the weight image in the pack buffer
  int16_t N, K;       // A-operand rows (output features) and contraction length
  int8_t layer, kind, net, in_buf, load_x, save, qdst, out_buf;
};

struct QDev {
  rb200_dqn_args_t a;
  rb200_net_ws_t ws;
  const unsigned char* pack;
  int nsteps, last_fwd_step;
  int buf_off[3];  // operand buffers (bytes from the smem base); [2] holds dZ of the last layer
  int q_off, ldq, lin_off, bar_off;
  int acc_cols, a_stages;  // tensor memory: accumulator columns, then a_stages weight stages
  int need_zero;    // some layer width is not a multiple of 8: clear the operand buffers first
  int dbg_mode;     // profiling only (bits): 1 skip the N=32 MMAs, 2 skip the N=64 MMAs, 4 no ring reads,
                    // 8 no bulk-copy traffic, 16 no tensor-memory stores, 32 no epilogue stores
  long long* dbg;  // optional timeline of block 0: [step][8] clock64 stamps (profiling builds)
  QStep steps[kQMaxSteps];
};

// ---------------------------------------------------------------------------
// weight packing: fp32 arena -> tiled chunk images
// ---------------------------------------------------------------------------
struct PackJob {
  const float* W;   // nn.Linear weight [rows_src x ld]
  int N, K;         // operand rows / contraction length (after the optional transpose)
  int ld;           // source row stride
  int transpose;    // 0: A[m][k] = W[m][k]; 1: A[m][k] = W[k][m]
  uint32_t pack_off;
  int chunk0;       // index of this job's first chunk in the grid
};
struct PackDev {
  PackJob jobs[3 * kMaxLayers];
  int njobs;
  unsigned char* pack;
};

constexpr int kPackParts = 4;  // blocks per chunk: one round of loads per thread at KC = 32

__global__ void __launch_bounds__(256) dqn_tc_pack_kernel(const PackDev p) {
  const int chunk = (int)blockIdx.x / kPackParts, part = (int)blockIdx.x % kPackParts;
  int j = 0;
  while (j + 1 < p.njobs && chunk >= p.jobs[j + 1].chunk0) ++j;
  const PackJob job = p.jobs[j];
  const int local = chunk - job.chunk0;
  const int kch = ceil_div(job.K, kQKC);
  const int t = local / kch, c = local - t * kch;
  const ChunkGeo g = chunk_geo(job.N, job.K, t, c);
  const int rows8 = (int)(g.lbo - 16) / 16, kl8 = g.ksteps * 8;
  float* img = reinterpret_cast<float*>(p.pack + job.pack_off + g.off);
  const int m0 = 128 * t, k0 = kQKC * c;
  const int total = rows8 * kl8;
  constexpr int U = 4;
  for (int base = (part * 256 + (int)threadIdx.x); base < total; base += kPackParts * 256 * U) {
    float v[U];
    int o[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * kPackParts * 256;
      int m, kk;
      if (job.transpose) { kk = idx / rows8; m = idx - kk * rows8; }
      else { m = idx / kl8; kk = idx - m * kl8; }
      v[u] = 0.f;
      o[u] = -1;
      if (idx < total) {
        o[u] = (kk >> 2) * (int)(g.lbo / 4) + m * 4 + (kk & 3);
        if (m0 + m < job.N && k0 + kk < job.K)
          v[u] = job.transpose ? __ldg(job.W + (size_t)(k0 + kk) * job.ld + m0 + m)
                               : __ldg(job.W + (size_t)(m0 + m) * job.ld + k0 + kk);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (o[u] >= 0) img[o[u]] = v[u];
  }
}

// ---------------------------------------------------------------------------
// the TD kernel
// ---------------------------------------------------------------------------
// activations other than ReLU / linear go through out-of-line calls so that the unrolled
// epilogues stay small enough for the instruction cache
__device__ __noinline__ float act_fwd_slow(float x, int act) { return act_fwd(x, act); }
__device__ __noinline__ float act_bwd_slow(float y, int act) { return act_bwd_from_out(y, act); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// two 16-column groups in flight, one wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t t0, uint32_t t1, uint32_t (&v)[16],
                                            uint32_t (&w)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(t0));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]),
        "=r"(w[7]), "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]),
        "=r"(w[14]), "=r"(w[15])
      : "r"(t1));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// 32 consecutive TMEM columns of this thread's lane <- 32 registers
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(taddr),
      "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
      "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]),
      "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]),
      "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
      "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
      "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}
// D[tmem] (+)= A[tmem: 128 lanes x 8 tf32 columns] . B[smem descriptor]
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Register cap: threads x registers must fit the 64 K register file (kQRegs; the kernel needs
// 76-107 depending on the cap, no spills).
__global__ void __maxnreg__(kQRegs)
dqn_td_tc_kernel(const Mlp q, const Mlp qt, const QDev p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const rb200_dqn_args_t& a = p.a;
  const int B = a.batch;
  const int row0 = blockIdx.x * kQR;
  const int L = q.n_layers;
  const int A = q.dims[L];
  auto gtime = []() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (long long)t; };
  if (kTimeline && p.dbg && tid == 0) p.dbg[kQMaxSteps * 8 + blockIdx.x * 4 + 0] = gtime();

  // operand padding (k up to the next multiple of 8) must be finite.  With every layer width a
  // multiple of 8 there is no padding: all operand quads, all 64 B-operand rows (rows past the
  // batch are stored as zeros) and the loss inputs are fully written before they are read, and
  // the ~115 KB clear (about 1 us per CTA) is skipped.
  if (p.need_zero) {
    unsigned nbytes;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(nbytes));
    // (the weight ring is fully overwritten by the bulk copies; rows a partial tile over-reads
    // only feed accumulator lanes nobody looks at)
    for (unsigned i = (unsigned)p.buf_off[0] + tid * 16u; i + 15u < nbytes; i += kQThreads * 16u)
      *reinterpret_cast<float4*>(smem_raw + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async_smem();
  }
  __syncthreads();

  unsigned char* ring = smem_raw;
  float* qarr = reinterpret_cast<float*>(smem_raw + p.q_off);  // [3][kQR][ldq]: q(s') online, target, q(s)
  float* act_s = reinterpret_cast<float*>(smem_raw + p.lin_off);  // [kQR][A] action weights
  float* mask_s = act_s + kQR * A;                                // [kQR][A] next-action mask
  float* scal_s = mask_s + kQR * A;                               // [kQR][4] reward, not_terminal, discount src
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + p.bar_off);
  uint64_t* sfree = full + kQStages;
  uint64_t* afull = sfree + kQStages;
  uint64_t* adone = afull + kAMaxStages;
  uint64_t* dready = adone + kAMaxStages;  // one per accumulator tile (see the epilogue)
  uint64_t* opready = dready + kQMaxTiles;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(opready + 1);
  const int ldq = p.ldq;

  if (tid == 0) {
    // a chunk is handled by the four warps of ONE loader group
    for (int s = 0; s < kQStages; ++s) { mbar_init(full + s, 1); mbar_init(sfree + s, 4); }
    for (int t = 0; t < kAMaxStages; ++t) { mbar_init(afull + t, 4); mbar_init(adone + t, 1); }
    for (int t = 0; t < kQMaxTiles; ++t) mbar_init(dready + t, 1);
    mbar_init(opready, kQEpiThreads / 32);  // one arrival per epilogue warp
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     smem_u32(tmem_slot)), "n"(kQTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (kTimeline && p.dbg && tid == 0) p.dbg[kQMaxSteps * 8 + blockIdx.x * 4 + 1] = gtime();

  // warp-uniform role index (the shuffle lets the compiler keep the role loops in uniform registers)
  const int role = __shfl_sync(0xffffffffu, warp, 0);
  if (role == kQEpiThreads / 32) {
    // =====================  weight producer warp (bulk copies)  =====================
    // Free-running over the static chunk list; the `sfree` barriers of the ring are the only
    // back-pressure, so up to kQStages chunks are in flight ahead of the loader warps.
    const bool leader = elect_one();
    int stage = 0;
    uint32_t par = 1;  // parity of the PREVIOUS use of `stage` (first lap: passes immediately)
    for (int s = 0; s < p.nsteps; ++s) {
      const QStep st = p.steps[s];
      const int mt = ceil_div(st.N, 128), kch = ceil_div(st.K, kQKC);
      const uint32_t tile_stride = (uint32_t)(round_up8(st.K) / 4) * kQFullLbo;
      for (int t = 0; t < mt; ++t) {
        const int rows = st.N - 128 * t;
        const uint32_t lbo = (uint32_t)(round_up8(rows < 128 ? rows : 128) * 16 + 16);
        const uint32_t full_bytes = (kQKC / 4) * lbo;
        const int klast = st.K - kQKC * (kch - 1);
        const uint32_t last_bytes = (uint32_t)(round_up8(klast) / 4) * lbo;
        const unsigned char* src = p.pack + st.pack_off + (size_t)t * tile_stride;
        for (int c = 0; c < kch; c += kQSub) {
          const int nsub = kch - c < kQSub ? kch - c : kQSub;
          const uint32_t bytes = (uint32_t)(nsub - 1) * full_bytes + ((c + nsub == kch) ? last_bytes : full_bytes);
          const uint32_t cbytes = (kTimeline && (p.dbg_mode & 8)) ? 16u : bytes;  // profiling: no copy traffic
          mbar_wait(sfree + stage, par);
          if (leader) {
            mbar_expect_tx(full + stage, cbytes);
            bulk_g2s(ring + stage * kQStageBytes, src, cbytes, full + stage);
          }
          src += bytes;
          if (++stage == kQStages) { stage = 0; par ^= 1u; }
        }
      }
    }
  } else if (role == kQEpiThreads / 32 + 1) {
    // =====================  MMA issuer warp  =====================
    // A operand: the split weights in tensor memory (128 lanes = features, 8 columns per
    // k step; hi columns then lo columns of the chunk).  B operand: the activations in
    // shared memory (descriptor).
    const bool leader = elect_one();
    const uint32_t idesc64 = umma_idesc_tf32(128, 64), idesc32 = umma_idesc_tf32(128, 32);
    const uint64_t desc_hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;  // SBO = 128 B, version 1
    int ts = 0;
    uint32_t tpar = 0;
    for (int s = 0; s < p.nsteps; ++s) {
      const QStep st = p.steps[s];
      const int mt = ceil_div(st.N, 128), kch = ceil_div(st.K, kQKC);
      mbar_wait(opready, (uint32_t)s & 1u);
      tc_fence_after();
      long long wafull = 0, tfence = 0, tissue = 0, tcommit = 0;
      if (kTimeline && p.dbg && blockIdx.x == 0 && leader) p.dbg[s * 8 + 0] = clock64();
      const uint32_t b0 = ((smem_u32(smem_raw + p.buf_off[st.in_buf]) >> 4) & 0x3fffu) |
                          ((uint32_t)(kQLboB >> 4) << 16);
      for (int t = 0; t < mt; ++t) {
        const uint32_t d = tmem + (uint32_t)(t * 64);
        uint32_t bdesc = b0;  // advances by two k quads per MMA k step
        for (int c = 0; c < kch; c += kQSub) {
          const int nsub = kch - c < kQSub ? kch - c : kQSub;
          const long long w0 = (kTimeline && p.dbg) ? clock64() : 0;
          mbar_wait(afull + ts, tpar);
          const long long w1 = (kTimeline && p.dbg) ? clock64() : 0;
          if (kTimeline && p.dbg) wafull += w1 - w0;
          tc_fence_after();
          const long long w2 = (kTimeline && p.dbg) ? clock64() : 0;
          if (kTimeline && p.dbg) tfence += w2 - w1;
          if (leader) {
            uint32_t bd = bdesc;
            for (int j = 0; j < nsub; ++j) {
              const int kl = st.K - kQKC * (c + j);
              const int ksteps = round_up8(kl < kQKC ? kl : kQKC) / 8;
              uint32_t a_hi = tmem + (uint32_t)(p.acc_cols + ts * kAStageCols + j * kASubCols);
              for (int ks = 0; ks < ksteps; ++ks) {
                if (!kTimeline || !(p.dbg_mode & 2))
                  umma_tf32_ts(d, a_hi, desc_hi | bd, idesc64, (c + j > 0 || ks > 0) ? 1u : 0u);
                if (!kTimeline || !(p.dbg_mode & 1))
                  umma_tf32_ts(d, a_hi + kQKC, desc_hi | bd, idesc32, 1u);
                a_hi += 8;
                bd += (2u * kQLboB) >> 4;
              }
            }
            const long long w3 = (kTimeline && p.dbg) ? clock64() : 0;
            umma_commit(adone + ts);  // this TMEM stage may be refilled once the MMAs retired
            if (kTimeline && p.dbg) { const long long w4 = clock64(); tissue += w3 - w2; tcommit += w4 - w3; }
          }
          bdesc += (uint32_t)nsub * (uint32_t)(kQKC / 4) * (kQLboB >> 4);
          if (++ts == p.a_stages) { ts = 0; tpar ^= 1u; }
        }
        // one accumulator tile complete: its epilogue runs while the next tile's MMAs issue
        if (leader) umma_commit(dready + t);
      }
      if (kTimeline && p.dbg && blockIdx.x == 0 && leader) {
        p.dbg[s * 8 + 1] = clock64();
        p.dbg[s * 8 + 2] = wafull;
        long long* fine = p.dbg + kQMaxSteps * 8 + 4 * 4096 + s * 8;
        fine[0] = tfence; fine[1] = tissue; fine[2] = tcommit;
      }
      __syncwarp();
    }
  } else if (role >= kQEpiThreads / 32 + 2) {
    // =====================  loader warps: shared memory -> TF32 split -> tensor memory  ======
    // warp (10 + j) may touch TMEM lanes [32*((10+j)%4), +32): the four loader warps cover the
    // four lane quadrants; lane i of a warp owns weight row 32*quadrant + i of the tile.
    const int quadrant = warp & 3;
    const int group = (warp - (kQEpiThreads / 32 + 2)) >> 2;
    int turn = 0;  // group whose chunk comes next; every warp walks the whole chunk sequence
    const int r = quadrant * 32 + lane;
    int ss = 0, ts = 0;
    uint32_t spar = 0, tpar = 1;  // TMEM stages start free (parity of the previous use)
    for (int s = 0; s < p.nsteps; ++s) {
      const QStep st = p.steps[s];
      const int mt = ceil_div(st.N, 128), kch = ceil_div(st.K, kQKC);
      long long lwfull = 0, lwdone = 0, lsplit = 0, lstore = 0;
      for (int t = 0; t < mt; ++t) {
        const int rows = st.N - 128 * t;
        const int rows8 = round_up8(rows < 128 ? rows : 128);
        const uint32_t lbo = (uint32_t)(rows8 * 16 + 16);
        for (int c = 0; c < kch; c += kQSub) {
          const int nsub = kch - c < kQSub ? kch - c : kQSub;
          const bool mine = turn == group;
          if (++turn == kQLoaderGroups) turn = 0;
          if (!mine) {
            if (++ss == kQStages) { ss = 0; spar ^= 1u; }
            if (++ts == p.a_stages) { ts = 0; tpar ^= 1u; }
            continue;
          }
          const long long l0 = (kTimeline && p.dbg) ? clock64() : 0;
          mbar_wait(full + ss, spar);
          const long long l0b = (kTimeline && p.dbg) ? clock64() : 0;
          if (kTimeline && p.dbg) lwfull += l0b - l0;
          const uint32_t ta = tmem + ((uint32_t)(quadrant * 32) << 16) +
                              (uint32_t)(p.acc_cols + ts * kAStageCols);
#pragma unroll 1
          for (int j = 0; j < nsub; ++j) {
            const int kl = st.K - kQKC * (c + j);
            const int nq = round_up8(kl < kQKC ? kl : kQKC) / 4;
            float hi[32], lo[32];
            const unsigned char* src = ring + ss * kQStageBytes + (uint32_t)j * (kQKC / 4) * lbo + r * 16;
#pragma unroll
            for (int qd = 0; qd < kQKC / 4; ++qd) {
              float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
              if (kTimeline && (p.dbg_mode & 4)) v = make_float4(1.f, 2.f, 3.f, 4.f);  // profiling: no ring reads
              else if (qd < nq && r < rows8) v = *reinterpret_cast<const float4*>(src + qd * lbo);
              float4 h, l;
              split4(v, h, l);
              hi[4 * qd + 0] = h.x; hi[4 * qd + 1] = h.y; hi[4 * qd + 2] = h.z; hi[4 * qd + 3] = h.w;
              lo[4 * qd + 0] = l.x; lo[4 * qd + 1] = l.y; lo[4 * qd + 2] = l.z; lo[4 * qd + 3] = l.w;
            }
            if (j == nsub - 1) {
              // the stage's values are in registers: the shared-memory stage can be refilled
              __syncwarp();
              if (lane == 0) mbar_arrive(sfree + ss);
            }
            if (j == 0) {
              const long long l1 = (kTimeline && p.dbg) ? clock64() : 0;
              if (kTimeline && p.dbg) lsplit += l1 - l0b;
              mbar_wait(adone + ts, tpar);  // the MMAs that read this TMEM stage have retired
              if (kTimeline && p.dbg) lwdone += clock64() - l1;
              tc_fence_after();
            }
            if (!kTimeline || !(p.dbg_mode & 16)) {  // (profiling: no tensor-memory stores)
              tmem_st32(ta + j * kASubCols, hi);
              tmem_st32(ta + j * kASubCols + kQKC, lo);
            }
          }
          const long long l2 = (kTimeline && p.dbg) ? clock64() : 0;
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(afull + ts);
          if (kTimeline && p.dbg) lstore += clock64() - l2;
          if (++ss == kQStages) { ss = 0; spar ^= 1u; }
          if (++ts == p.a_stages) { ts = 0; tpar ^= 1u; }
        }
      }
      if (kTimeline && p.dbg && blockIdx.x == 0 && quadrant == 0 && group == 0 && lane == 0) {
        p.dbg[s * 8 + 6] = lwfull;
        p.dbg[s * 8 + 7] = lwdone;
        long long* fine = p.dbg + kQMaxSteps * 8 + 4 * 4096 + s * 8;
        fine[3] = lsplit; fine[4] = lstore;
      }
    }
  } else {
    // =====================  operand producers / epilogue warps  =====================
    const int quad = warp & 3, grp = warp >> 2;
    uint32_t dphase = 0;       // bit t: parity of the next phase of dready[t]

    // The input tile of a pass: global -> registers (x_fetch, issued early so that the load
    // latency hides behind the previous layer) -> hi/lo split -> B operand (x_store).
    constexpr int kXQ = 4;  // float4 pieces per thread held in flight (covers S <= 128)
    const int xS = q.dims[0];
    const int xnq = round_up8(xS) / 4;
    const bool x_in_regs = kQR * xnq <= kXQ * kQEpiThreads;
    float4 xr[kXQ];
    auto x_fetch_one = [&](const float* src, int idx) {
      const bool vec = ((xS & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
      const int r = idx / xnq, qd = idx - r * xnq;
      const int k = 4 * qd, row = row0 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < kQR * xnq && row < B) {
        const float* sp = src + (size_t)row * xS;
        if (vec && k + 3 < xS) {
          // volatile: keep the load HERE (the compiler would sink it next to its use)
          asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];\n"
                       : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                       : "l"(sp + k));
        } else {
          if (k < xS) v.x = sp[k];
          if (k + 1 < xS) v.y = sp[k + 1];
          if (k + 2 < xS) v.z = sp[k + 2];
          if (k + 3 < xS) v.w = sp[k + 3];
        }
      }
      return v;
    };
    auto x_store_one = [&](int idx, const float4 v) {
      if (idx >= kQR * xnq) return;
      float* base = reinterpret_cast<float*>(smem_raw + p.buf_off[0]);
      const int r = idx / xnq, qd = idx - r * xnq;
      float4 h, l;
      split4(v, h, l);
      const int o = qd * (kQLboB / 4) + r * 4;
      *reinterpret_cast<float4*>(base + o) = h;
      *reinterpret_cast<float4*>(base + o + kQLoOff) = l;
    };
    auto x_fetch = [&](const float* src) {
      if (!x_in_regs) return;
#pragma unroll
      for (int i = 0; i < kXQ; ++i) xr[i] = x_fetch_one(src, tid + i * kQEpiThreads);
    };
    auto x_store = [&](const float* src) {
      if (x_in_regs) {
#pragma unroll
        for (int i = 0; i < kXQ; ++i) x_store_one(tid + i * kQEpiThreads, xr[i]);
      } else {
        for (int idx = tid; idx < kQR * xnq; idx += kQEpiThreads) x_store_one(idx, x_fetch_one(src, idx));
      }
    };
    // per-row inputs of the loss, staged while the first layers run
    auto load_loss_inputs = [&]() {
      const float* mask = a.maxq ? a.possible_next_actions_mask : a.next_action;
      for (int idx = tid; idx < kQR * A; idx += kQEpiThreads) {
        const int r = idx / A;
        const bool in = row0 + r < B;
        const size_t g = (size_t)row0 * A + idx;
        act_s[idx] = in ? a.action[g] : 0.f;
        mask_s[idx] = (in && mask) ? mask[g] : 1.f;
      }
      if (tid < kQR) {
        const int row = row0 + tid;
        const bool in = row < B;
        scal_s[tid * 4 + 0] = in ? a.reward[row] : 0.f;
        scal_s[tid * 4 + 1] = in ? a.not_terminal[row] : 0.f;
        scal_s[tid * 4 + 2] = (in && a.discount_mode == RB200_DISCOUNT_POW) ? a.discount_src[row] : 0.f;
      }
    };
    auto epilogue = [&](const QStep& st) {
      const Mlp& net = st.net ? qt : q;
      const int l = st.layer, N = st.N;
      const int mt = ceil_div(N, 128);
      float* obase = reinterpret_cast<float*>(smem_raw + p.buf_off[st.out_buf]);
      for (int t = 0; t < mt; ++t) {
        // One barrier per tile: the MMA warp runs ahead through the tiles of a step without
        // waiting for the epilogues, so a shared barrier could complete two phases before a
        // slow thread looks at it (and parity waits cannot tell "two ahead" from "not yet").
        // Per tile there is at most one completion per step, and steps are serialised by
        // `opready`.
        mbar_wait(dready + t, (dphase >> t) & 1u);
        dphase ^= 1u << t;
        tc_fence_after();
        const int h16 = grp * 16;
        const int n = t * 128 + quad * 32 + lane;
        const bool valid = n < N;
        uint32_t v[16], w[16];
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * 64 + h16);
        tmem_ld16x2(taddr, taddr + 32, v, w);
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(v[j]) + __uint_as_float(w[j]);
        float* ob = obase + (n >> 2) * (kQLboB / 4) + h16 * 4 + (n & 3);
        bool to_operand, to_global;
        float* gdst = nullptr;
        if (st.kind == kStepBwd) {
          // act'(h_{l-1}) from the forward operand still resident in shared memory (h = hi + lo
          // exactly); dZ_{l-1} then replaces it in place as the next B operand
          const int hact = q.act[l - 1];
          // st.save: the shared-memory copy of h_{l-1} was overwritten by a later forward layer
          // (networks with >= 3 hidden layers ping-pong over the same two buffers); read the
          // copy saved for the weight-gradient kernel instead
          float hv[16];
          if (st.save) {
            const float* hs = p.ws.hidden[l - 1];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int row = row0 + h16 + j;
              hv[j] = (valid && row < B) ? hs[(size_t)row * N + n] : 0.f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) hv[j] = valid ? ob[j * 4] + ob[j * 4 + kQLoOff] : 0.f;
          }
          if (hact == RB200_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = hv[j] > 0.f ? x[j] : 0.f;
          } else if (hact != RB200_ACT_LINEAR) {
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = valid ? x[j] * act_bwd_slow(hv[j], hact) : 0.f;
          }
          to_operand = l - 1 >= 1;
          to_global = true;
          gdst = p.ws.dz[l - 1];
        } else {
          const float bias = valid ? __ldg(net.params + net.b_off[l] + n) : 0.f;
          const int act = net.act[l];
          if (act == RB200_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = fmaxf(x[j] + bias, 0.f);
          } else if (act == RB200_ACT_LINEAR) {
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] += bias;
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = act_fwd_slow(x[j] + bias, act);
          }
          to_operand = st.kind != kStepLast;
          to_global = st.save != 0 && st.kind != kStepLast;
          gdst = to_global ? p.ws.hidden[l] : nullptr;
        }
        if (!valid) continue;
        if (kTimeline && (p.dbg_mode & 32)) { to_operand = false; to_global = false; }  // profiling: no epilogue stores
        if (st.kind == kStepLast) {
          float* qd = qarr + (st.qdst * kQR + h16) * ldq + n;
#pragma unroll
          for (int j = 0; j < 16; ++j) qd[j * ldq] = x[j];
          continue;
        }
        const int nrow = B - (row0 + h16);  // rows of this half that exist
        if (to_operand) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float h, lo_;
            split1(j < nrow ? x[j] : 0.f, h, lo_);
            ob[j * 4] = h;
            ob[j * 4 + kQLoOff] = lo_;
          }
        }
        if (to_global) {
          float* gd = gdst + (size_t)(row0 + h16) * N + n;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < nrow) gd[(size_t)j * N] = x[j];
        }
      }
    };
    auto loss_stage = [&]() {
      // same arithmetic as rb200_dqn.cu (dqn_trainer.py:157-239); 8 lanes per batch row, the
      // actions strided over them, combined with shuffles inside the 8-lane group
      const float* qa = qarr;
      const float* qb = qarr + kQR * ldq;
      const float* qc = qarr + 2 * kQR * ldq;
      const int r = tid >> 3, sub = tid & 7;
      const int row = row0 + r;
      const bool in = row < B;
      // arg max over the (masked) next-state values: first index wins ties, as a sequential
      // "key > best" scan does
      float best = 0.f, sel = 0.f, qsel = 0.f, bsum = 0.f;
      int bi = 0x7fffffff;
      for (int c = sub; c < A; c += 8) {
        const float pen = -1e9f * (1.f - mask_s[r * A + c]);
        const float vt = qb[r * ldq + c] + pen;
        const float key = a.double_q ? (qa[r * ldq + c] + pen) : vt;
        if (bi == 0x7fffffff || key > best) { best = key; bi = c; sel = vt; }
        const float aw = act_s[r * A + c];
        qsel += qc[r * ldq + c] * aw;
        if (a.reward_boost) bsum += aw * a.reward_boost[c];
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const float os = __shfl_xor_sync(0xffffffffu, sel, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        qsel += __shfl_xor_sync(0xffffffffu, qsel, o);
        bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
        if (oi != 0x7fffffff && (bi == 0x7fffffff || ob > best || (ob == best && oi < bi))) {
          best = ob; bi = oi; sel = os;
        }
      }
      float le = 0.f, g = 0.f;
      if (in) {
        const float rew = scal_s[r * 4 + 0] + bsum;
        const float disc = (a.discount_mode == RB200_DISCOUNT_POW)
                               ? powf(a.gamma, scal_s[r * 4 + 2]) : a.gamma;
        const float tgt = rew + disc * (sel * scal_s[r * 4 + 1]);
        const float d = qsel - tgt;
        const float invB = 1.f / (float)B;
        if (a.loss_kind == RB200_LOSS_HUBER) {
          const float ad = fabsf(d);
          le = ad < 1.f ? 0.5f * d * d : ad - 0.5f;
          g = (d < -1.f) ? -invB : (d > 1.f ? invB : invB * d);
        } else {
          le = d * d;
          g = 2.f * invB * d;
        }
        if (sub == 0) {
          if (a.td_target) a.td_target[row] = tgt;
          if (a.next_action_idx) a.next_action_idx[row] = bi;
          if (a.q_selected) a.q_selected[row] = qsel;
        }
      }
      const int A8 = round_up8(A);
      const int lact = q.act[L - 1];
      float* zb = reinterpret_cast<float*>(smem_raw + p.buf_off[2]) + r * 4;
      for (int c = sub; c < A8; c += 8) {
        float v = 0.f;
        if (in && c < A) {
          v = g * act_s[r * A + c];
          if (lact != RB200_ACT_LINEAR) v *= act_bwd_slow(qc[r * ldq + c], lact);
          if (a.all_action_scores) a.all_action_scores[(size_t)row * A + c] = qc[r * ldq + c];
          if (a.do_backward) p.ws.dz[L - 1][(size_t)row * A + c] = v;
        }
        if (a.do_backward) {
          float h, lo_;
          split1(v, h, lo_);
          float* o = zb + (c >> 2) * (kQLboB / 4) + (c & 3);
          o[0] = h;
          o[kQLoOff] = lo_;
        }
      }
      const float ws = warp_sum(sub == 0 ? le : 0.f);
      if (lane == 0) scal_s[kQR * 4 + warp] = ws;  // per-warp loss sums, combined by thread 0 at the end
    };

    auto x_src = [&](int lx) { return lx == 1 ? a.state : a.next_state; };
    if (p.steps[0].load_x) { x_fetch(x_src(p.steps[0].load_x)); x_store(x_src(p.steps[0].load_x)); }
    // hand-over: every thread makes its operand stores visible to the async proxy, the warp
    // converges, ONE lane arrives (8 arrivals per hand-over instead of 256 serialised ones)
    auto operand_ready = [&]() {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(opready);
    };
    operand_ready();
    load_loss_inputs();
    for (int s = 0; s < p.nsteps; ++s) {
      const QStep st = p.steps[s];
      const int lx = (s + 1 < p.nsteps) ? p.steps[s + 1].load_x : 0;
      if (lx) x_fetch(x_src(lx));  // in flight during this step's MMAs and epilogue
      if (kTimeline && p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[s * 8 + 3] = clock64();
      epilogue(st);
      tc_fence_before();
      if (kTimeline && p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[s * 8 + 4] = clock64();
      if (s == p.last_fwd_step) {
        asm volatile("bar.sync 1, %0;\n" ::"n"(kQEpiThreads) : "memory");
        loss_stage();
      }
      if (s + 1 < p.nsteps) {
        if (lx) x_store(x_src(lx));
        operand_ready();
        if (kTimeline && p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[s * 8 + 5] = clock64();
      }
    }

    // publish the loss (off the critical path of the step loop): last tile reduces
    asm volatile("bar.sync 1, %0;\n" ::"n"(kQEpiThreads) : "memory");
    if (tid == 0) {
      float loss_partial = 0.f;
      for (int w = 0; w < kQEpiThreads / 32; ++w) loss_partial += scal_s[kQR * 4 + w];
      a.loss_partials[blockIdx.x] = loss_partial;
      __threadfence();
      const unsigned fin = atomicAdd(a.tile_counter, 1u);
      if (fin == gridDim.x - 1) {
        __threadfence();
        float tot = 0.f;
        for (unsigned i = 0; i < gridDim.x; ++i) tot += ((volatile float*)a.loss_partials)[i];
        *a.loss = tot / (float)B;
        *a.tile_counter = 0u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kTimeline && p.dbg && tid == 0) p.dbg[kQMaxSteps * 8 + blockIdx.x * 4 + 2] = gtime();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem),
                 "n"(kQTmemCols));
  }
}

// ---------------------------------------------------------------------------
// host side: plan (steps, pack jobs, shared-memory layout)
// ---------------------------------------------------------------------------
struct QPlan {
  QDev dev;
  PackDev pack;
  int pack_chunks;
  size_t smem_bytes;
  int64_t pack_bytes;
  bool ok;
};

static QPlan make_plan(const rb200_mlp_t* qn, const rb200_mlp_t* qtn, int double_q, int do_backward) {
  QPlan pl;
  memset(&pl, 0, sizeof(pl));
  pl.ok = false;
  const int L = qn->n_layers;
  if (L < 1 || L > kMaxLayers) return pl;
  for (int l = 1; l <= L; ++l)
    if (qn->dims[l] > 128 * kQMaxTiles || qn->dims[l] > 32000) return pl;
  if (qn->dims[0] > 32000 || qn->dims[L] > 256) return pl;

  // weight images: online fwd, target fwd, online bwd (transposed)
  uint32_t off = 0;
  uint32_t off_on[kMaxLayers], off_tg[kMaxLayers], off_bw[kMaxLayers];
  int nj = 0, nchunks = 0;
  auto add_job = [&](const rb200_mlp_t* net, int l, int transpose) {
    PackJob& j = pl.pack.jobs[nj++];
    j.W = net->params + net->w_off[l];
    j.ld = net->dims[l];
    j.transpose = transpose;
    j.N = transpose ? net->dims[l] : net->dims[l + 1];
    j.K = transpose ? net->dims[l + 1] : net->dims[l];
    j.pack_off = off;
    j.chunk0 = nchunks;
    nchunks += ceil_div(j.N, 128) * ceil_div(j.K, kQKC);
    const uint32_t o = off;
    off += image_bytes(j.N, j.K);
    return o;
  };
  for (int l = 0; l < L; ++l) off_on[l] = add_job(qn, l, 0);
  if (qtn) for (int l = 0; l < L; ++l) off_tg[l] = add_job(qtn, l, 0);
  else for (int l = 0; l < L; ++l) { off_tg[l] = off; off += image_bytes(qn->dims[l + 1], qn->dims[l]); nj++; }
  if (do_backward) for (int l = 1; l < L; ++l) off_bw[l] = add_job(qn, l, 1);
  pl.pack.njobs = nj;
  pl.pack_chunks = nchunks;
  {
    // the Adam kernel writes the same images from rb200_dqn_tc_layout.cuh's table
    const TcImages im = tc_images(qn, do_backward);
    pl.pack_bytes = im.total_bytes;
    bool same = im.total_bytes == (int64_t)off + 4096;
    for (int l = 0; l < L; ++l) same = same && im.on_fwd[l] == off_on[l] && im.tg_fwd[l] == off_tg[l];
    if (do_backward) for (int l = 1; l < L; ++l) same = same && im.on_bwd[l] == off_bw[l];
    if (!same) return pl;  // ok stays false: layout tables disagree (a bug, not a user error)
  }

  // steps
  int ns = 0;
  auto add_pass = [&](int net, int load_x, int save, int qdst) {
    for (int l = 0; l < L; ++l) {
      QStep& s = pl.dev.steps[ns++];
      s.pack_off = net ? off_tg[l] : off_on[l];
      s.N = (int16_t)qn->dims[l + 1];
      s.K = (int16_t)qn->dims[l];
      s.layer = (int8_t)l;
      s.kind = (l == L - 1) ? kStepLast : kStepHidden;
      s.net = (int8_t)net;
      s.in_buf = (int8_t)(l & 1);
      s.out_buf = (int8_t)((l + 1) & 1);
      s.load_x = (int8_t)(l == 0 ? load_x : 0);
      s.save = (int8_t)save;
      s.qdst = (int8_t)qdst;
    }
  };
  add_pass(1, 2, 0, 1);                    // q_target(next_state)
  if (double_q) add_pass(0, 2, 0, 0);      // q(next_state)
  add_pass(0, 1, do_backward ? 1 : 0, 2);  // q(state)
  pl.dev.last_fwd_step = ns - 1;
  if (do_backward) {
    for (int l = L - 1; l >= 1; --l) {
      QStep& s = pl.dev.steps[ns++];
      s.pack_off = off_bw[l];
      s.N = (int16_t)qn->dims[l];
      s.K = (int16_t)qn->dims[l + 1];
      s.layer = (int8_t)l;
      s.kind = kStepBwd;
      s.net = 0;
      // dZ of the last layer sits in its own buffer so that the forward operands h_{l-1}
      // (the activation derivatives) survive until their backward step; dZ_{l-1} then
      // overwrites h_{l-1} in place
      s.in_buf = (int8_t)(l == L - 1 ? 2 : ((l + 1) & 1));
      s.out_buf = (int8_t)(l & 1);
      s.save = (int8_t)(l <= L - 3 ? 1 : 0);  // h_{l-1} clobbered by h_{l+1}: use the global copy
    }
  }
  pl.dev.nsteps = ns;
  {
    // tensor memory: accumulators for the widest layer, every other column is weight ring
    int tiles = 1;
    for (int l = 1; l <= L; ++l) tiles = tiles > ceil_div(qn->dims[l], 128) ? tiles : ceil_div(qn->dims[l], 128);
    pl.dev.acc_cols = 64 * tiles;
    int st = (kQTmemCols - pl.dev.acc_cols) / kAStageCols;
    st = st < kAMaxStages ? st : kAMaxStages;
    // a ring stage must belong to ONE loader group (tests/test_tc_protocol_model.py): depth is a
    // multiple of the group count
    pl.dev.a_stages = st / kQLoaderGroups * kQLoaderGroups;
    if (pl.dev.a_stages < 2) return pl;
  }
  pl.dev.need_zero = 0;
  for (int l = 0; l <= L; ++l)
    if (qn->dims[l] % 8 != 0) pl.dev.need_zero = 1;

  // shared memory
  int maxd[3] = {8, 8, qn->dims[L]};
  for (int i = 0; i < L; ++i) if (qn->dims[i] > maxd[i & 1]) maxd[i & 1] = qn->dims[i];
  size_t o = (size_t)kQStages * kQStageBytes;
  for (int b = 0; b < 3; ++b) {
    pl.dev.buf_off[b] = (int)o;
    o += (size_t)(round_up8(maxd[b]) / 4) * kQLboB;
  }
  pl.dev.ldq = qn->dims[L] + 1;
  pl.dev.q_off = (int)o;
  o += (size_t)3 * kQR * pl.dev.ldq * sizeof(float);
  o = (o + 15) & ~(size_t)15;
  pl.dev.lin_off = (int)o;
  o += ((size_t)2 * kQR * qn->dims[L] + 4 * kQR + 8) * sizeof(float);
  o = (o + 15) & ~(size_t)15;
  pl.dev.bar_off = (int)o;
  o += (2 * kQStages + 2 * kAMaxStages + kQMaxTiles + 1) * sizeof(uint64_t) + 16;
  pl.smem_bytes = (o + 15) & ~(size_t)15;
  pl.ok = pl.smem_bytes <= (size_t)kQMaxSmem;
  return pl;
}

}  // namespace rb200

using namespace rb200;

static long long* g_tc_dbg = nullptr;
// profiling hook (not part of the reference-facing API): device buffer of kQMaxSteps*8 int64
extern "C" void rb200_debug_set_tc_timeline(void* dev_buf) { g_tc_dbg = static_cast<long long*>(dev_buf); }

extern "C" int64_t rb200_dqn_tc_workspace_bytes(const rb200_mlp_t* q_net, int32_t double_q,
                                                int32_t do_backward) {
  if (!q_net || validate_mlp(q_net, "q_network")) return 0;
  const QPlan pl = make_plan(q_net, nullptr, double_q, do_backward);
  return pl.ok ? pl.pack_bytes : 0;
}

extern "C" int rb200_dqn_tc_pack(const rb200_mlp_t* q_net, const rb200_mlp_t* q_target,
                                 int32_t double_q, int32_t do_backward, void* pack_ws,
                                 int64_t pack_ws_bytes, void* stream) {
  if (!q_net || !q_target || !pack_ws) { set_last_error("rb200_dqn_tc_pack: null argument"); return RB200_E_INVALID; }
  if (int rc = validate_mlp(q_net, "q_network")) return rc;
  if (int rc = validate_mlp(q_target, "q_network_target")) return rc;
  if (q_net->n_layers != q_target->n_layers) { set_last_error("q_network / target layer count mismatch"); return RB200_E_INVALID; }
  for (int l = 0; l <= q_net->n_layers; ++l)
    if (q_net->dims[l] != q_target->dims[l]) { set_last_error("q_network / target dims mismatch at %d", l); return RB200_E_INVALID; }
  if ((reinterpret_cast<uintptr_t>(pack_ws) & 127) != 0) { set_last_error("pack workspace must be 128-byte aligned"); return RB200_E_INVALID; }
  QPlan pl = make_plan(q_net, q_target, double_q, do_backward);
  if (!pl.ok) { set_last_error("rb200_dqn_tc_pack: shapes do not fit the tcgen05 path"); return RB200_E_SMEM; }
  if (pack_ws_bytes < pl.pack_bytes) { set_last_error("pack workspace too small: %lld < %lld", (long long)pack_ws_bytes, (long long)pl.pack_bytes); return RB200_E_INVALID; }
  pl.pack.pack = static_cast<unsigned char*>(pack_ws);
  dqn_tc_pack_kernel<<<pl.pack_chunks * kPackParts, 256, 0, (cudaStream_t)stream>>>(pl.pack);
  return check_cuda(cudaGetLastError(), "dqn_tc_pack_kernel launch");
}

extern "C" int rb200_dqn_td_step_tc(const rb200_mlp_t* q_net, const rb200_mlp_t* q_target,
                                    const rb200_dqn_args_t* args, const rb200_net_ws_t* ws,
                                    void* pack_ws, int64_t pack_ws_bytes, int32_t weights_packed,
                                    void* stream) {
  if (!q_net || !q_target || !args || !ws || !pack_ws) { set_last_error("rb200_dqn_td_step_tc: null argument"); return RB200_E_INVALID; }
  if (!weights_packed) {
    if (int rc = rb200_dqn_tc_pack(q_net, q_target, args->double_q, args->do_backward, pack_ws, pack_ws_bytes, stream)) return rc;
  }
  if (int rc = validate_mlp(q_net, "q_network")) return rc;
  if (int rc = validate_mlp(q_target, "q_network_target")) return rc;
  if (q_net->n_layers != q_target->n_layers) { set_last_error("q_network / target layer count mismatch"); return RB200_E_INVALID; }
  for (int l = 0; l <= q_net->n_layers; ++l)
    if (q_net->dims[l] != q_target->dims[l]) { set_last_error("q_network / target dims mismatch at %d", l); return RB200_E_INVALID; }
  if (args->batch <= 0) { set_last_error("batch must be positive"); return RB200_E_INVALID; }
  if (!args->state || !args->next_state || !args->action || !args->reward || !args->not_terminal ||
      !args->loss_partials || !args->loss || !args->tile_counter) {
    set_last_error("rb200_dqn_td_step_tc: required pointer is null"); return RB200_E_INVALID;
  }
  if (!args->maxq && !args->next_action) { set_last_error("SARSA update needs next_action"); return RB200_E_INVALID; }
  if (args->discount_mode == RB200_DISCOUNT_POW && !args->discount_src) { set_last_error("POW discount needs discount_src"); return RB200_E_INVALID; }
  if (args->do_backward) {
    for (int l = 0; l < q_net->n_layers; ++l)
      if (!ws->dz[l] || (l < q_net->n_layers - 1 && !ws->hidden[l])) { set_last_error("workspace buffer missing for layer %d", l); return RB200_E_INVALID; }
  }
  if ((reinterpret_cast<uintptr_t>(pack_ws) & 127) != 0) { set_last_error("pack workspace must be 128-byte aligned"); return RB200_E_INVALID; }
  QPlan pl = make_plan(q_net, q_target, args->double_q, args->do_backward);
  if (!pl.ok) { set_last_error("rb200_dqn_td_step_tc: shapes do not fit the tcgen05 path"); return RB200_E_SMEM; }
  if (pack_ws_bytes < pl.pack_bytes) { set_last_error("pack workspace too small: %lld < %lld", (long long)pack_ws_bytes, (long long)pl.pack_bytes); return RB200_E_INVALID; }
  pl.dev.a = *args;
  pl.dev.ws = *ws;
  pl.dev.pack = static_cast<const unsigned char*>(pack_ws);
  pl.dev.dbg = g_tc_dbg;
  { const char* e = getenv("RB200_TC_DBG_MODE"); pl.dev.dbg_mode = e ? atoi(e) : 0; }
  cudaStream_t st = (cudaStream_t)stream;
  static SmemOptIn optin = {};  // per device; raised outside graph capture by the first eager call
  {
    cudaError_t e = ensure_dynamic_smem(dqn_td_tc_kernel, optin, pl.smem_bytes);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(dqn_td_tc)");
  }
  const Mlp q = make_mlp(q_net), qt = make_mlp(q_target);
  const int grid = ceil_div(args->batch, kQR);
  dqn_td_tc_kernel<<<grid, kQThreads, pl.smem_bytes, st>>>(q, qt, pl.dev);
  return check_cuda(cudaGetLastError(), "dqn_td_tc_kernel launch");
}
