// pe_mlp.cu — fused positional-bias MLP of GroupedVectorAttention (the "next" row).
//
// Replaces the torch chain of
// /root/reference/pointcept/models/point_transformer_v2/point_transformer_v2m2_base.py:88-93,116-118
//     peb = linear_p_bias(pos)  =  Linear(C,C)( ReLU( PointBatchNorm(C)( Linear(3,C)(pos) ) ) )
// on all N·k neighbour rows — five passes over (N,k,C) tensors forward (Linear out, BN read+write, ReLU,
// Linear out, dtype copies) and about ten backward — by
//   forward : read pos (12 B/row), write peb (4C B/row)
//   backward: read grad_peb (4C B/row) + pos, write only parameter gradients.
//
// What makes the fusion possible with TRAINING-mode BatchNorm: the first layer is affine in pos, so its
// batch statistics follow from the mean and covariance of pos in closed form,
//     mean_c = W1[c,:]·m + b1[c],   var_c = W1[c,:] · Cov · W1[c,:]ᵀ        (biased, as BN uses)
// and pos (hence m, Cov) is shared by every block of a BlockSequence.  With a_c = γ_c·rstd_c the hidden
// activation is  h[r,c] = relu(a_c·(W1[c,:]·pos[r]) + a_c·(b1[c]-mean_c) + β_c): three FMAs per element
// recomputed on the fly, never stored.  The C×C layer runs on the 5th-generation tensor cores (tcgen05.mma,
// bf16 operands, fp32 accumulation in Tensor Memory — the precision of the reference's autocast path); the work
// is HBM-bound (2·C flops per output byte).  (Round 1's mma.sync kernels — forward 43 %, backward 17 % of the HBM
// peak, bound by the shared-memory pipe that fed the MMA — were measured against these and removed:
// profiles/r02a_kernel_bench_count_csr_mma_pe.txt, r02c_kernel_bench_pe_tc.txt.)
//
// The backward pass needs only row sums (∂L/∂pos is not needed: coordinates are inputs):
//     dz = (G·W2) ⊙ [z>0],   S1 = Σ dz⊗pos,  S2 = Σ dz,  S3 = Σ dz⊙x̂,   dW2 = Gᵀ·h,  db2 = Σ G
//     dγ = S3, dβ = S2, dW1 = rstd ⊙ [γ S1 − (γ S2/R)⊗Σpos − (γ S3/R) ⊙ rstd ⊙ (W1·(Σ pos posᵀ − Σpos Σposᵀ/R))]
//     db1 = 0 (a bias in front of a training-mode BatchNorm has no gradient)
// One pass over G: per-CTA partial sums (fixed order, no atomics) + a small finalize kernel.
// Supported widths: C in {48, 96} (levels 0 and 1 of the PTv2m2 configs: 76 % of the (N,k,C) traffic);
// other widths keep the torch path (ao_b200/ptv2.py).
#include <cuda_bf16.h>

#include "common.cuh"

namespace aopt {

constexpr int kMomBlock = 256;

// ---- 1. moments of pos ------------------------------------------------------------------------------
// out[0..2] = Σp, out[3..8] = Σ xx, xy, xz, yy, yz, zz (double).  Two stages, fixed summation order.
__global__ void __launch_bounds__(kMomBlock)
pos_moments_partial_kernel(long long rows, const float *__restrict__ pos, double *__restrict__ partial) {
    double acc[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) acc[i] = 0.0;
    const long long step = (long long)gridDim.x * kMomBlock;
    for (long long r = (long long)blockIdx.x * kMomBlock + threadIdx.x; r < rows; r += step) {
        const float x = __ldg(pos + r * 3), y = __ldg(pos + r * 3 + 1), z = __ldg(pos + r * 3 + 2);
        acc[0] += x; acc[1] += y; acc[2] += z;
        acc[3] += (double)x * x; acc[4] += (double)x * y; acc[5] += (double)x * z;
        acc[6] += (double)y * y; acc[7] += (double)y * z; acc[8] += (double)z * z;
    }
    __shared__ double red[9][kMomBlock / 32];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        double v = acc[i];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) == 0) red[i][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        double v = 0.0;
        for (int w = 0; w < kMomBlock / 32; ++w) v += red[threadIdx.x][w];
        partial[(size_t)blockIdx.x * 9 + threadIdx.x] = v;
    }
}

// One warp per moment (9 warps): lanes stride over the per-CTA partials, then a shuffle tree — a fixed order (one thread
// per moment walking all 592 partials took 37 us: profiles/r02c_ops_L0_ncu_summary.md).
__global__ void __launch_bounds__(288)
pos_moments_final_kernel(int n_partial, const double *__restrict__ partial, double *__restrict__ out) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double v = 0.0;
    for (int i = lane; i < n_partial; i += 32) v += partial[(size_t)i * 9 + w];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) out[w] = v;
}

// ---- 2. fold BatchNorm into per-channel affine maps, convert W2 ---------------------------------------
// fold[c] = { a·W1[c,0..2], a·(b1-mean)+β,  rstd·W1[c,0..2], rstd·(b1-mean) }   (z = fold[0..3]·(p,1), x̂ = fold[4..7]·(p,1))
// stats[c] = { mean, biased var, rstd }.  use_batch = 1: statistics from the moments; 0: running stats.
__global__ void __launch_bounds__(128)
pe_fold_kernel(int c, double rows, const double *__restrict__ mom, const float *__restrict__ w1,
               const float *__restrict__ b1, const float *__restrict__ gamma, const float *__restrict__ beta,
               const float *__restrict__ running_mean, const float *__restrict__ running_var, float eps,
               int use_batch, const float *__restrict__ w2, float *__restrict__ fold, float *__restrict__ stats,
               __nv_bfloat16 *__restrict__ w2_bf, __nv_bfloat16 *__restrict__ w2t_bf,
               const float *__restrict__ aux_w, int ga, __nv_bfloat16 *__restrict__ wf_bf,
               __nv_bfloat16 *__restrict__ wft_bf) {
    const int ch = blockIdx.x;
    if (threadIdx.x == 0) {
        const double wx = w1[ch * 3], wy = w1[ch * 3 + 1], wz = w1[ch * 3 + 2];
        double mean, var;
        if (use_batch) {
            const double mx = mom[0] / rows, my = mom[1] / rows, mz = mom[2] / rows;
            const double cxx = mom[3] / rows - mx * mx, cxy = mom[4] / rows - mx * my, cxz = mom[5] / rows - mx * mz;
            const double cyy = mom[6] / rows - my * my, cyz = mom[7] / rows - my * mz, czz = mom[8] / rows - mz * mz;
            mean = wx * mx + wy * my + wz * mz + (double)b1[ch];
            var = wx * (cxx * wx + cxy * wy + cxz * wz) + wy * (cxy * wx + cyy * wy + cyz * wz) +
                  wz * (cxz * wx + cyz * wy + czz * wz);
            if (var < 0.0) var = 0.0;
        } else {
            mean = running_mean[ch];
            var = running_var[ch];
        }
        const double rstd = 1.0 / sqrt(var + (double)eps);
        const double a = (double)gamma[ch] * rstd;
        const double shift = (double)b1[ch] - mean;
        float *f = fold + (size_t)ch * 8;
        f[0] = (float)(a * wx); f[1] = (float)(a * wy); f[2] = (float)(a * wz); f[3] = (float)(a * shift + (double)beta[ch]);
        f[4] = (float)(rstd * wx); f[5] = (float)(rstd * wy); f[6] = (float)(rstd * wz); f[7] = (float)(rstd * shift);
        stats[ch] = (float)mean;
        stats[c + ch] = (float)var;
        stats[2 * c + ch] = (float)rstd;
    }
    for (int i = threadIdx.x; i < c; i += 128) {  // row ch of W2 [co][ci] and column ch of its transpose
        const __nv_bfloat16 v = __float2bfloat16_rn(w2[(size_t)ch * c + i]);
        w2_bf[(size_t)ch * c + i] = v;
        w2t_bf[(size_t)i * c + ch] = v;
    }
    // auxiliary head (ga <= 16 extra outputs): wf [16][c] (rows >= ga zero) and its transpose wft [c][16]
    if (threadIdx.x < 16) {
        const int gq = threadIdx.x;
        const float v = (aux_w && gq < ga) ? aux_w[(size_t)gq * c + ch] : 0.f;
        wf_bf[(size_t)gq * c + ch] = __float2bfloat16_rn(v);
        wft_bf[(size_t)ch * 16 + gq] = __float2bfloat16_rn(v);
    }
}

// ---- bulk asynchronous copies (TMA 1-D: cp.async.bulk → SASS UBLKCP) ---------------------------------------
// A tile of TR consecutive rows is one contiguous TR·C·4-byte block in global memory, so it moves with a
// single bulk copy issued by one thread: no LSU wavefronts, no registers, completion through an mbarrier
// (loads) or a bulk group (stores).
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---- 3b. forward on the 5th-generation tensor cores (tcgen05.mma, accumulator in Tensor Memory) -----------
// One CTA = 128 threads = one 128-row tile at a time (UMMA M = 128: accumulator row i lives in TMEM lane i).
//   A = h tile (128 x C bf16), written by the threads themselves: thread r recomputes the hidden activations of row
//       r from pos (three FMAs + ReLU per element) and stores them as 16-byte pieces in the canonical K-major,
//       no-swizzle operand layout  byte(row, k) = (k/8)·2048 + row·16 + (k%8)·2  (8x8 "core matrices": SBO = 128 B
//       between 8-row groups, LBO = 2048 B between the two 8-wide K halves of one K=16 step).  Consecutive threads
//       write consecutive 16-byte words: conflict-free, and no ldmatrix / LSU traffic to feed the MMA afterwards —
//       the mma.sync kernel above spent its time in exactly that (shared-memory pipe bound).
//   B = [W2 ; W_aux] (NPAD x C bf16, K-major, same layout with LBO = NPAD·16), resident in shared memory.
//   D = 128 x NPAD fp32 in TMEM (NPAD columns), C/16 MMAs of K = 16 issued by ONE thread, completion through
//       tcgen05.commit -> mbarrier.  Epilogue: every warp reads its 32 lanes with tcgen05.ld.32x32b (thread = row),
//       adds b2 and streams the row to HBM with 128-bit stores.
// Phases of a tile are serial inside a CTA (produce A | MMA | epilogue); several CTAs per SM (TMEM columns and
// shared memory allow 8 at C=48, 4 at C=96, 1-2 at C=192) overlap them.  The work is HBM-bound: 4·(C+ga)+12 B/row.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start address [0,14) >>4 | LBO [16,30) >>4 | SBO [32,46) >>4 | version = 1 at [46,48)
    // | base offset 0 | layout type [61,64) = 0 (SWIZZLE_NONE)
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__device__ __forceinline__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
    // cute::UMMA::InstrDescriptor (kind::f16): D = f32 (bits [4,6) = 1), A = B = bf16 ([7,10) = [10,13) = 1), both
    // K-major (bits 15, 16 = 0), N >> 3 at [17,23), M >> 4 at [24,29)
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

constexpr int kTcBlock = 128;   // threads = rows of a tile = TMEM lanes

template <int C>
struct PeTcLayout {
    static constexpr int NPAD = C + 16;                         // [W2 ; 16 auxiliary rows]: a multiple of 16 (UMMA M=128)
    static constexpr int TMEM_COLS = NPAD <= 64 ? 64 : NPAD <= 128 ? 128 : NPAD <= 256 ? 256 : 512;
    static constexpr uint32_t A_LBO = kTcBlock * 16, B_LBO = NPAD * 16, SBO = 128;
    static constexpr size_t a_bytes = (size_t)kTcBlock * C * 2, b_bytes = (size_t)NPAD * C * 2;
    static constexpr int OUT_LD = C + 4;                         // floats per staged output row: 16 bytes of padding
    static constexpr size_t out_bytes = (size_t)kTcBlock * OUT_LD * 4;
    // the staged output rows re-use the A tile's memory (A is dead once the tile's MMAs have completed)
    static constexpr size_t ao_bytes = a_bytes > out_bytes ? a_bytes : out_bytes;
    static constexpr size_t bytes = ao_bytes + b_bytes + 16 * (size_t)C + 4 * (size_t)C + 64;
    static constexpr int ctas_per_sm = (512 / TMEM_COLS) < (int)(232448 / (bytes + 1024)) ? (512 / TMEM_COLS) : (int)(232448 / (bytes + 1024));
};

template <int C>
__global__ void __launch_bounds__(kTcBlock)
pe_mlp_forward_tc_kernel(long long rows, const float *__restrict__ pos, const float *__restrict__ fold,
                         const __nv_bfloat16 *__restrict__ w2_bf, const float *__restrict__ b2,
                         float *__restrict__ out, const __nv_bfloat16 *__restrict__ wf_bf, int ga,
                         float *__restrict__ aux_out) {
    using L = PeTcLayout<C>;
    constexpr int NPAD = L::NPAD;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *a_s = smem_raw;                                             // A: [C/8][128][8] bf16 ...
    float *out_s = reinterpret_cast<float *>(smem_raw);                       // ... and, after the MMAs, [128][C+4] staged output rows
    unsigned char *b_s = a_s + L::ao_bytes;                                    // B: [C/8][NPAD][8] bf16
    float4 *fz = reinterpret_cast<float4 *>(b_s + L::b_bytes);                 // [C] folded z-map
    float *b2s = reinterpret_cast<float *>(fz + C);                            // [C]
    uint64_t *bar = reinterpret_cast<uint64_t *>(b2s + C);                     // MMA-complete barrier
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool aux = aux_out != nullptr;

    // B operand: rows 0..C-1 = W2[co][:], rows C..C+15 = the auxiliary head (rows >= ga are zero), 16 bytes a piece
    for (int i = tid; i < NPAD * (C / 8); i += kTcBlock) {
        const int n = i % NPAD, k8 = i / NPAD;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (n < C) v = *reinterpret_cast<const uint4 *>(w2_bf + (size_t)n * C + k8 * 8);
        else if (aux) v = *reinterpret_cast<const uint4 *>(wf_bf + (size_t)(n - C) * C + k8 * 8);
        *reinterpret_cast<uint4 *>(b_s + (size_t)k8 * L::B_LBO + n * 16) = v;
    }
    for (int i = tid; i < C; i += kTcBlock) {
        fz[i] = make_float4(fold[i * 8], fold[i * 8 + 1], fold[i * 8 + 2], fold[i * 8 + 3]);
        b2s[i] = b2[i];
    }
    if (tid == 0) mbar_init(bar, 1);
    if (warp == 0) {  // one warp allocates the accumulator columns and gives the permit back
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(L::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a_addr = smem_u32(a_s), b_addr = smem_u32(b_s);
    constexpr uint32_t idesc = umma_idesc_bf16(128, NPAD);
    const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);      // this warp's 32 lanes

    const long long n_tiles = (rows + kTcBlock - 1) / kTcBlock;
    uint32_t phase = 0;
    // position of this thread's row in the NEXT tile, requested one tile ahead (its latency used to sit at the head of
    // every tile, in front of the whole serial produce -> MMA -> epilogue chain)
    float nx = 0.f, ny = 0.f, nz = 0.f;
    {
        const long long g0 = (long long)blockIdx.x * kTcBlock + tid;
        if (blockIdx.x < n_tiles && g0 < rows) { nx = __ldg(pos + g0 * 3); ny = __ldg(pos + g0 * 3 + 1); nz = __ldg(pos + g0 * 3 + 2); }
    }
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long gr = tile * kTcBlock + tid;
        const bool live = gr < rows;
        // ---- A: hidden activations of this thread's row, 8 channels (16 bytes) at a time ----
        const float px = nx, py = ny, pz = nz;
        {
            const long long gn = (tile + gridDim.x) * kTcBlock + tid;
            nx = ny = nz = 0.f;
            if (tile + gridDim.x < n_tiles && gn < rows) { nx = __ldg(pos + gn * 3); ny = __ldg(pos + gn * 3 + 1); nz = __ldg(pos + gn * 3 + 2); }
        }
        // the A tile shares its memory with the staged output rows of the previous tile: every thread's bulk copy must
        // have finished READING before anybody writes A (each thread waits for its own copy, the barrier below joins them)
        bulk_store_wait_read();
        __syncthreads();
#pragma unroll
        for (int k8 = 0; k8 < C / 8; ++k8) {
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 f0 = fz[k8 * 8 + 2 * q], f1 = fz[k8 * 8 + 2 * q + 1];
                const float z0 = fmaf(f0.x, px, fmaf(f0.y, py, fmaf(f0.z, pz, f0.w)));
                const float z1 = fmaf(f1.x, px, fmaf(f1.y, py, fmaf(f1.z, pz, f1.w)));
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(live ? fmaxf(z0, 0.f) : 0.f, live ? fmaxf(z1, 0.f) : 0.f);
                w[q] = *reinterpret_cast<const uint32_t *>(&h2);
            }
            *reinterpret_cast<uint4 *>(a_s + (size_t)k8 * L::A_LBO + tid * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_proxy_async();   // generic-proxy writes of A -> visible to the tensor core (async proxy)
        tc_fence_before();     // orders this thread's tcgen05.ld of the previous tile before the barrier
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < C / 16; ++kk) {
                const uint64_t da = umma_smem_desc(a_addr + kk * 2 * L::A_LBO, L::A_LBO, L::SBO);
                const uint64_t db = umma_smem_desc(b_addr + kk * 2 * L::B_LBO, L::B_LBO, L::SBO);
                umma_bf16(tmem_base, da, db, idesc, kk > 0 ? 1u : 0u);
            }
            umma_commit(bar);  // arrives when the MMAs above have completed (implies fence::before_thread_sync)
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();
        // ---- epilogue: TMEM lane = row; 16 columns per load.  The row goes to this thread's OWN padded slot in shared
        // memory (row stride C+4 floats: the 16-byte stores of a quarter warp hit distinct banks) and from there to HBM
        // with one bulk copy per row (cp.async.bulk, 4C contiguous bytes).  Storing straight from registers made every
        // warp-wide store touch 32 different 128-byte lines (row stride 4C): 453 us at level 0 against 371 us for the
        // mma.sync kernel, whose tile left through a bulk copy too (profiles/r02a_kernel_bench.txt).  No barrier: the slot
        // is private to the thread, which waits for its previous copy to have READ the slot before rewriting it. ----
        float *srow = out_s + tid * L::OUT_LD;
#pragma unroll
        for (int j = 0; j < C / 16; ++j) {
            float v[16];
            tmem_ld16(tmem_row + j * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int col = j * 16 + q * 4;
                *reinterpret_cast<float4 *>(srow + col) = make_float4(v[q * 4] + b2s[col], v[q * 4 + 1] + b2s[col + 1],
                                                                       v[q * 4 + 2] + b2s[col + 2], v[q * 4 + 3] + b2s[col + 3]);
            }
        }
        fence_proxy_async();   // this thread's generic-proxy writes -> visible to the bulk-copy (async) proxy
        if (live) bulk_store(out + gr * C, srow, (uint32_t)(C * 4));
        if (aux) {
            float v[16];
            tmem_ld16(tmem_row + C, v);
            tmem_ld_wait();
            if (live) {
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    if (q < ga) aux_out[gr * ga + q] = v[q];
            }
        }
    }
    bulk_store_wait_all();   // shared memory must outlive this thread's last copy
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(L::TMEM_COLS) : "memory");
}

// ---- 4b. backward on tcgen05: three UMMA products per 128-row tile, parameter-gradient accumulators resident in TMEM ----
// Operand tiles live in shared memory as 16-byte pieces  byte(chunk, row) = chunk·2048 + row·16  (chunk = 8 consecutive
// channels, row = 0..127).  ONE such buffer is two canonical no-swizzle UMMA layouts at once:
//   K-major  (M/N = row,     K = channel): SBO = 128 (8-row groups),     LBO = 2048 (8-channel groups)
//   MN-major (M/N = channel, K = row)    : SBO = 2048 (8-channel groups), LBO = 128  (8-row groups)
// so the gradient tile GU = [G | dU] feeds  dh = GU · [W2 ; Wf]  as a K-major A operand and, untouched, the transposed
// product  P = GUᵀ · [h | 1]  (= dW2, db2, dWf) as an MN-major A operand — no transposed copy (the mma.sync kernel kept
// G and h in both orientations and was bound by the shared-memory pipe: 17 % of the HBM roofline).
//   (1) D1 (128 x C)        = GU (128 x C+16, K-major)  · W2T ((C+16) x C, K-major B)        fresh every tile
//   (2) dz = D1 ⊙ [h > 0] -> bf16 into the GU buffer (free once (1) and (3) have completed)
//   (3) P  ((C+16) x (C+16)) += GUᵀ (MN-major A, K = 128 rows) · [h | 1 | 0] (MN-major B)       accumulates over tiles
//   (4) Q  (C x 16)          += dzᵀ (MN-major A)              · [p_hi | 1 | p_lo | 0] (MN-major B)   accumulates over tiles
// P and Q stay in Tensor Memory for the whole life of the CTA; they are read once at the end into the per-CTA partial
// sums the finalize kernels already consume.  S3 = Σ dz ⊙ x̂ is linear in (S1, S2): x̂ = fx·(p, 1), so it needs no sum of
// its own.  pos enters (4) as bf16 hi + lo parts (16 significant bits: dW1 subtracts Σ dz ⊗ mean(p) from S1).
// UMMA M = 64 accumulators (C = 48: 64 channel rows) sit in lanes 0-15 of each 32-lane quarter: row m <-> lane
// (m % 16) + 32 (m / 16)  (cute::UMMA::tmem_frg, 1-SM, M = 64); M = 128: row m <-> lane m.
__device__ __forceinline__ constexpr uint32_t umma_idesc_bf16_mn(int m, int n) {   // both operands MN-major
    return umma_idesc_bf16(m, n) | (1u << 15) | (1u << 16);
}

template <int C>
struct PeTcBwd {
    static constexpr int CA = C + 16;                          // channels of [G | dU]  (aux gradient padded to 16)
    static constexpr int NP = C + 16;                          // columns of P: [h | 1 | 0 ...]
    static constexpr int MP = CA <= 64 ? 64 : 128;             // UMMA M of P (rows = channels of GU)
    static constexpr int MQ = C <= 64 ? 64 : 128;              // UMMA M of Q (rows = channels of dz)
    static constexpr int GU_CHUNKS = (MP > MQ ? MP : MQ) / 8;  // the GU buffer is re-used for dz
    static constexpr int H_CHUNKS = NP / 8;
    // bytes from one chunk (8 channels x 128 rows) to the next: 2048 + 16.  The 16 bytes of padding rotate the banks of
    // successive chunks, so that the 16-byte pieces the gradient load writes for ONE row (6 chunks, same row) no longer
    // fall on the same four banks (19 M conflict wavefronts of 66 M at level 0, LSU pipe 79 %: profiles/r02c); the
    // UMMA descriptors take the stride as it is (LBO of the K-major view, SBO of the MN-major view).
    static constexpr uint32_t CHUNK = kTcBlock * 16 + 16;
    static constexpr uint32_t W_LBO = C * 16;                  // W2T: [(C+16)/8][C][8], K-major B of product (1)
    static constexpr size_t gu_bytes = (size_t)GU_CHUNKS * CHUNK, h_bytes = (size_t)H_CHUNKS * CHUNK, pp_bytes = 2 * (size_t)CHUNK,
                            w_bytes = (size_t)(CA / 8) * W_LBO;
    static constexpr size_t bytes = gu_bytes + h_bytes + pp_bytes + w_bytes + 16 * (size_t)C + 64;
    static constexpr int COL_D1 = 0, COL_P = C, COL_Q = C + NP;
    static constexpr int TMEM_COLS = (2 * C + 32) <= 128 ? 128 : 256;
    static constexpr int ctas_per_sm = (512 / TMEM_COLS) < (int)(232448 / (bytes + 1024)) ? (512 / TMEM_COLS) : (int)(232448 / (bytes + 1024));
    static constexpr int partial_floats = C * C + 6 * C + 16 * C;   // per-CTA partial sums: dW2 [C*C] | db2 [C] | S2 [C] | S3 [C] | S1 [C][3] | dWf [16][C]
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&v);
}

template <int C>
__global__ void __launch_bounds__(kTcBlock)
pe_mlp_backward_tc_kernel(long long rows, const float *__restrict__ pos, const float *__restrict__ fold,
                          const __nv_bfloat16 *__restrict__ w2t_bf, const float *__restrict__ grad,
                          float *__restrict__ partial, const __nv_bfloat16 *__restrict__ wft_bf, int ga,
                          const float *__restrict__ grad_aux) {
    using L = PeTcBwd<C>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *gu_s = smem_raw;                                   // [G | dU] tile, later the dz tile
    unsigned char *h_s = gu_s + L::gu_bytes;                          // [h | 1 | 0] tile
    unsigned char *pp_s = h_s + L::h_bytes;                           // [p_hi, 1, p_lo, 0 | 0] tile
    unsigned char *w_s = pp_s + L::pp_bytes;                          // W2T: n = ci, k = co then the 16 aux rows
    float4 *fz = reinterpret_cast<float4 *>(w_s + L::w_bytes);        // [C] folded z-map
    uint64_t *bar = reinterpret_cast<uint64_t *>(fz + C);             // bar[0]: products (1)+(3) done, bar[1]: (4) done
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool aux = grad_aux != nullptr;

    // ---- one-time setup: W2T, the constant chunks of the h / pos tiles, barriers, TMEM ----
    for (int i = tid; i < C * (L::CA / 8); i += kTcBlock) {
        const int n = i % C, k8 = i / C;     // B(n = ci, k = 8 k8 ..): W2[co][ci] for co < C, Wf[g'][ci] for the last 16
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (k8 < C / 8) v = *reinterpret_cast<const uint4 *>(w2t_bf + (size_t)n * C + k8 * 8);
        else if (aux) v = *reinterpret_cast<const uint4 *>(wft_bf + (size_t)n * 16 + (k8 - C / 8) * 8);
        *reinterpret_cast<uint4 *>(w_s + (size_t)k8 * L::W_LBO + n * 16) = v;
    }
    for (int i = tid; i < C; i += kTcBlock) fz[i] = make_float4(fold[i * 8], fold[i * 8 + 1], fold[i * 8 + 2], fold[i * 8 + 3]);
    {   // column C of [h | 1 | 0] is the constant 1 (db2 = Σ G), the rest of the last two chunks is 0; second pos chunk 0
        const uint32_t one = pack_bf16x2(1.f, 0.f);
        *reinterpret_cast<uint4 *>(h_s + (size_t)(C / 8) * L::CHUNK + tid * 16) = make_uint4(one, 0u, 0u, 0u);
        *reinterpret_cast<uint4 *>(h_s + (size_t)(C / 8 + 1) * L::CHUNK + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4 *>(pp_s + L::CHUNK + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
        // chunks of the GU buffer beyond the real channels are only ever read as M-padding (ignored output rows)
        for (int ch = L::CA / 8; ch < L::GU_CHUNKS; ++ch)
            *reinterpret_cast<uint4 *>(gu_s + (size_t)ch * L::CHUNK + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(L::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t gu_addr = smem_u32(gu_s), h_addr = smem_u32(h_s), pp_addr = smem_u32(pp_s), w_addr = smem_u32(w_s);
    constexpr uint32_t idesc1 = umma_idesc_bf16(128, C);
    constexpr uint32_t idescP = umma_idesc_bf16_mn(L::MP, L::NP);
    constexpr uint32_t idescQ = umma_idesc_bf16_mn(L::MQ, 16);
    const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);

    const long long n_tiles = (rows + kTcBlock - 1) / kTcBlock;
    uint32_t phase = 0;
    bool first = true;
    // The gradient tile of the NEXT iteration, in registers: C/4 coalesced 128-bit loads per lane, issued right after this
    // tile's MMAs so that their DRAM latency runs under the epilogue instead of at the head of the next tile
    // (long-scoreboard was the top stall: profiles/r02c_ops_L0_ncu_summary.md row 21).
    float4 gnext[C / 4];
    auto load_tile = [&](long long t) {
        constexpr int F4 = C / 4;
        const long long wrow0 = t * kTcBlock + warp * 32;
        const float4 *src = reinterpret_cast<const float4 *>(grad + wrow0 * C);
#pragma unroll
        for (int i = 0; i < F4; ++i) {
            const int q = i * 32 + lane;
            const int rl = q / F4;
            gnext[i] = (t < n_tiles && wrow0 + rl < rows) ? ldg_stream4(reinterpret_cast<const float *>(src + q)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    load_tile(blockIdx.x);
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row_base = tile * kTcBlock;
        // ---- a. gradient tile: the warp's 32 rows are 32·C contiguous floats — coalesced 128-bit loads (issued ONE TILE
        //         AHEAD into registers, see the end of the loop body), two lanes assemble one 16-byte bf16 piece ----
        {
            constexpr int F4 = C / 4;                                 // float4 per row (even)
#pragma unroll
            for (int i = 0; i < F4; ++i) {
                const int q = i * 32 + lane;
                const int rl = q / F4, c4 = q - rl * F4;
                const float4 v = gnext[i];
                const uint32_t u0 = pack_bf16x2(v.x, v.y), u1 = pack_bf16x2(v.z, v.w);
                const uint32_t n0 = __shfl_down_sync(0xffffffffu, u0, 1), n1 = __shfl_down_sync(0xffffffffu, u1, 1);
                if ((lane & 1) == 0)
                    *reinterpret_cast<uint4 *>(gu_s + (size_t)(c4 >> 1) * L::CHUNK + (warp * 32 + rl) * 16) = make_uint4(u0, u1, n0, n1);
            }
        }
        const long long gr = row_base + tid;
        const bool live = gr < rows;
        {   // aux gradient of this thread's row: 16 channels (two chunks), zero beyond ga
            float du[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) du[q] = (aux && live && q < ga) ? __ldg(grad_aux + gr * ga + q) : 0.f;
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf)
                *reinterpret_cast<uint4 *>(gu_s + (size_t)(C / 8 + hlf) * L::CHUNK + tid * 16) =
                    make_uint4(pack_bf16x2(du[hlf * 8], du[hlf * 8 + 1]), pack_bf16x2(du[hlf * 8 + 2], du[hlf * 8 + 3]),
                               pack_bf16x2(du[hlf * 8 + 4], du[hlf * 8 + 5]), pack_bf16x2(du[hlf * 8 + 6], du[hlf * 8 + 7]));
        }
        // ---- b. hidden activations of this thread's row (as in the forward kernel) and its position ----
        float px = 0.f, py = 0.f, pz = 0.f;
        if (live) { px = __ldg(pos + gr * 3); py = __ldg(pos + gr * 3 + 1); pz = __ldg(pos + gr * 3 + 2); }
#pragma unroll
        for (int k8 = 0; k8 < C / 8; ++k8) {
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 f0 = fz[k8 * 8 + 2 * q], f1 = fz[k8 * 8 + 2 * q + 1];
                const float z0 = fmaf(f0.x, px, fmaf(f0.y, py, fmaf(f0.z, pz, f0.w)));
                const float z1 = fmaf(f1.x, px, fmaf(f1.y, py, fmaf(f1.z, pz, f1.w)));
                w[q] = pack_bf16x2(live ? fmaxf(z0, 0.f) : 0.f, live ? fmaxf(z1, 0.f) : 0.f);
            }
            *reinterpret_cast<uint4 *>(h_s + (size_t)k8 * L::CHUNK + tid * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        {
            const float hx = __bfloat162float(__float2bfloat16_rn(px)), hy = __bfloat162float(__float2bfloat16_rn(py)),
                        hz = __bfloat162float(__float2bfloat16_rn(pz));
            *reinterpret_cast<uint4 *>(pp_s + tid * 16) =
                make_uint4(pack_bf16x2(hx, hy), pack_bf16x2(hz, live ? 1.f : 0.f), pack_bf16x2(px - hx, py - hy), pack_bf16x2(pz - hz, 0.f));
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            // (1) dh = GU · W2T                                   M = 128 rows, N = C, K = C + 16
#pragma unroll
            for (int kk = 0; kk < L::CA / 16; ++kk)
                umma_bf16(tmem_base + L::COL_D1, umma_smem_desc(gu_addr + kk * 2 * L::CHUNK, L::CHUNK, 128),
                          umma_smem_desc(w_addr + kk * 2 * L::W_LBO, L::W_LBO, 128), idesc1, kk > 0 ? 1u : 0u);
            // (3) P += GUᵀ · [h | 1 | 0]                           M = channels of GU, N = C + 16, K = 128 rows
#pragma unroll
            for (int kk = 0; kk < kTcBlock / 16; ++kk)
                umma_bf16(tmem_base + L::COL_P, umma_smem_desc(gu_addr + kk * 256, 128, L::CHUNK),
                          umma_smem_desc(h_addr + kk * 256, 128, L::CHUNK), idescP, (!first || kk > 0) ? 1u : 0u);
            umma_commit(bar);
        }
        load_tile(tile + gridDim.x);   // next tile's gradient rows: in flight during the rest of this tile
        mbar_wait(bar, phase);
        tc_fence_after();
        // ---- c. dz = dh ⊙ [h > 0] of this thread's row -> bf16 pieces into the (now free) GU buffer ----
#pragma unroll
        for (int j = 0; j < C / 16; ++j) {
            float v[16];
            tmem_ld16(tmem_row + L::COL_D1 + j * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
                const uint4 hb = *reinterpret_cast<const uint4 *>(h_s + (size_t)(2 * j + hlf) * L::CHUNK + tid * 16);
                const uint32_t hw[4] = {hb.x, hb.y, hb.z, hb.w};
                uint32_t w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float d0 = (hw[q] & 0x7fffu) ? v[hlf * 8 + 2 * q] : 0.f;
                    const float d1 = (hw[q] & 0x7fff0000u) ? v[hlf * 8 + 2 * q + 1] : 0.f;
                    w[q] = pack_bf16x2(d0, d1);
                }
                *reinterpret_cast<uint4 *>(gu_s + (size_t)(2 * j + hlf) * L::CHUNK + tid * 16) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            // (4) Q += dzᵀ · [p_hi, 1, p_lo, 0]                    M = channels, N = 16, K = 128 rows
#pragma unroll
            for (int kk = 0; kk < kTcBlock / 16; ++kk)
                umma_bf16(tmem_base + L::COL_Q, umma_smem_desc(gu_addr + kk * 256, 128, L::CHUNK),
                          umma_smem_desc(pp_addr + kk * 256, 128, L::CHUNK), idescQ, (!first || kk > 0) ? 1u : 0u);
            umma_commit(bar + 1);
        }
        mbar_wait(bar + 1, phase);   // the tiles are rewritten by the next iteration
        tc_fence_after();
        phase ^= 1u;
        first = false;
    }
    // ---- d. per-CTA partial sums out of Tensor Memory (dW2 | db2 | S2 | S3 | S1 | dWf) ----
    float *out = partial + (size_t)blockIdx.x * L::partial_floats;
    if (first) {   // a CTA that owned no tile (cannot happen with pe_grid, kept for safety): zeros
        for (int i = tid; i < L::partial_floats; i += kTcBlock) out[i] = 0.f;
    } else {
        {   // P: row m = channel of GU
            const int m = L::MP == 64 ? warp * 16 + lane : warp * 32 + lane;
            const bool owner = (L::MP == 64 ? lane < 16 : true) && m < L::CA;
#pragma unroll
            for (int j = 0; j < L::NP / 16; ++j) {
                float v[16];
                tmem_ld16(tmem_row + L::COL_P + j * 16, v);
                tmem_ld_wait();
                if (owner) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const int n = j * 16 + q;
                        if (m < C) {
                            if (n < C) out[(size_t)m * C + n] = v[q];                     // dW2[co][ci]
                            else if (n == C) out[(size_t)C * C + m] = v[q];               // db2[co]
                        } else if (n < C) {
                            out[(size_t)C * C + 6 * C + (size_t)(m - C) * C + n] = v[q];  // dWf[g'][ci]
                        }
                    }
                }
            }
        }
        {   // Q: row m = channel of dz; columns p_hi (0..2), 1 (3), p_lo (4..6)
            const int m = L::MQ == 64 ? warp * 16 + lane : warp * 32 + lane;
            const bool owner = (L::MQ == 64 ? lane < 16 : true) && m < C;
            float v[16];
            tmem_ld16(tmem_row + L::COL_Q, v);
            tmem_ld_wait();
            if (owner) {
                const float s1x = v[0] + v[4], s1y = v[1] + v[5], s1z = v[2] + v[6], s2 = v[3];
                const float4 fx = make_float4(fold[m * 8 + 4], fold[m * 8 + 5], fold[m * 8 + 6], fold[m * 8 + 7]);
                out[(size_t)C * C + C + m] = s2;                                                   // S2
                out[(size_t)C * C + 2 * C + m] = fmaf(fx.x, s1x, fmaf(fx.y, s1y, fmaf(fx.z, s1z, fx.w * s2)));   // S3 = Σ dz·x̂
                out[(size_t)C * C + 3 * C + m * 3 + 0] = s1x;
                out[(size_t)C * C + 3 * C + m * 3 + 1] = s1y;
                out[(size_t)C * C + 3 * C + m * 3 + 2] = s1z;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(L::TMEM_COLS) : "memory");
}

// Sums the per-CTA partials in CTA order: dW2 directly, the channel sums (db2 | S2 | S3 | S1) into `sums`.
// 256 threads = 32 output elements x 8 slices of the partials; the slices are combined through shared memory in a
// fixed order (one thread per element walking all ~592 partials: 57 us per call, profiles/r02c_ops_L0_ncu_summary.md).
__global__ void __launch_bounds__(256)
pe_mlp_backward_finalize_kernel(int c, int n_partial, int partial_floats, const float *__restrict__ partial,
                                float *__restrict__ grad_w2, float *__restrict__ sums, int ga,
                                float *__restrict__ grad_aux_w) {
    __shared__ float red[8][32];
    const int total = c * c + 6 * c + (grad_aux_w ? ga * c : 0);
    const int e = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + e;
    float v = 0.f;
    if (i < total)
        for (int p = sl; p < n_partial; p += 8) v += partial[(size_t)p * partial_floats + i];
    red[sl][e] = v;
    __syncthreads();
    if (sl == 0 && i < total) {
        float t = red[0][e];
#pragma unroll
        for (int q = 1; q < 8; ++q) t += red[q][e];
        if (i < c * c) grad_w2[i] = t;
        else if (i < c * c + 6 * c) sums[i - c * c] = t;
        else grad_aux_w[i - c * c - 6 * c] = t;  // dWf [ga][c]
    }
}

// BatchNorm / Linear1 backward algebra on the channel sums (see the file header).
__global__ void __launch_bounds__(128)
pe_mlp_backward_params_kernel(int c, double rows, int use_batch, const float *__restrict__ sums,
                              const double *__restrict__ mom, const float *__restrict__ w1,
                              const float *__restrict__ gamma, const float *__restrict__ stats,
                              float *__restrict__ grad_w1, float *__restrict__ grad_b1,
                              float *__restrict__ grad_gamma, float *__restrict__ grad_beta,
                              float *__restrict__ grad_b2) {
    const int ch = blockIdx.x * 128 + threadIdx.x;
    if (ch >= c) return;
    const double db2 = sums[ch], s2 = sums[c + ch], s3 = sums[2 * c + ch];
    const double s1[3] = {sums[3 * c + ch * 3], sums[3 * c + ch * 3 + 1], sums[3 * c + ch * 3 + 2]};
    const double gam = gamma[ch], rstd = stats[2 * c + ch];
    grad_b2[ch] = (float)db2;
    grad_gamma[ch] = (float)s3;
    grad_beta[ch] = (float)s2;
    if (use_batch) {
        const double wx = w1[ch * 3], wy = w1[ch * 3 + 1], wz = w1[ch * 3 + 2];
        const double sp[3] = {mom[0], mom[1], mom[2]};
        // R·Cov = Σ p pᵀ − Σp Σpᵀ / R
        const double cxx = mom[3] - sp[0] * sp[0] / rows, cxy = mom[4] - sp[0] * sp[1] / rows, cxz = mom[5] - sp[0] * sp[2] / rows;
        const double cyy = mom[6] - sp[1] * sp[1] / rows, cyz = mom[7] - sp[1] * sp[2] / rows, czz = mom[8] - sp[2] * sp[2] / rows;
        const double xp[3] = {rstd * (wx * cxx + wy * cxy + wz * cxz), rstd * (wx * cxy + wy * cyy + wz * cyz),
                              rstd * (wx * cxz + wy * cyz + wz * czz)};  // Σ_r x̂ ⊗ p
        for (int d = 0; d < 3; ++d)
            grad_w1[ch * 3 + d] = (float)(rstd * (gam * s1[d] - (gam * s2 / rows) * sp[d] - (gam * s3 / rows) * xp[d]));
        grad_b1[ch] = 0.f;
    } else {
        for (int d = 0; d < 3; ++d) grad_w1[ch * 3 + d] = (float)(rstd * gam * s1[d]);
        grad_b1[ch] = (float)(rstd * gam * s2);
    }
}

// Persistent grids: as many CTAs as fit per SM (shared memory / registers), at most one per tile.
static int pe_grid(long long rows, int tile_rows, int ctas_per_sm) {
    long long tiles = (rows + tile_rows - 1) / tile_rows;
    long long cap = (long long)kNumSM * ctas_per_sm;
    return (int)(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
}
static int pe_bwd_grid(long long rows, int c) {
    return c <= 48 ? pe_grid(rows, kTcBlock, PeTcBwd<48>::ctas_per_sm) : pe_grid(rows, kTcBlock, PeTcBwd<96>::ctas_per_sm);
}


}  // namespace aopt

using namespace aopt;

static size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int aopt_pe_mlp_supported(int c) { return (c == 48 || c == 96) ? 1 : 0; }

/* moments: 9 doubles (Σp, Σ xx xy xz yy yz zz). */
extern "C" size_t aopt_pos_moments_workspace_bytes(void) { return a256((size_t)kNumSM * 4 * 9 * sizeof(double)); }

extern "C" int aopt_pos_moments(int64_t rows, const float *pos, double *moments, void *workspace,
                                size_t workspace_bytes, aopt_stream_t stream) {
    if (rows < 0 || !moments) return AOPT_ERR_INVALID_ARGUMENT;
    if (rows > 0 && !pos) return AOPT_ERR_INVALID_ARGUMENT;
    if (!workspace || workspace_bytes < aopt_pos_moments_workspace_bytes()) return AOPT_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    const int grid = stride_grid(rows > 0 ? rows : 1, kMomBlock, 4);
    double *partial = static_cast<double *>(workspace);
    pos_moments_partial_kernel<<<grid, kMomBlock, 0, st>>>(rows, pos, partial);
    pos_moments_final_kernel<<<1, 288, 0, st>>>(grid, partial, moments);
    return check_launch(2);
}

/* Scratch that lives from forward to backward of one call (caller-allocated):
 * fold (8C floats) | stats (3C floats) | W2 bf16 (C*C) | W2ᵀ bf16 (C*C). */
extern "C" size_t aopt_pe_mlp_state_bytes(int c) {
    return a256(4 * (size_t)8 * c) + a256(4 * (size_t)3 * c) + 2 * a256(2 * (size_t)c * c) + 2 * a256(2 * (size_t)16 * c);
}

namespace {
struct PeState {
    float *fold, *stats;
    __nv_bfloat16 *w2, *w2t, *wf, *wft;
};
PeState carve_state(void *state, int c) {
    char *p = static_cast<char *>(state);
    PeState s;
    s.fold = reinterpret_cast<float *>(p); p += a256(4 * (size_t)8 * c);
    s.stats = reinterpret_cast<float *>(p); p += a256(4 * (size_t)3 * c);
    s.w2 = reinterpret_cast<__nv_bfloat16 *>(p); p += a256(2 * (size_t)c * c);
    s.w2t = reinterpret_cast<__nv_bfloat16 *>(p); p += a256(2 * (size_t)c * c);
    s.wf = reinterpret_cast<__nv_bfloat16 *>(p); p += a256(2 * (size_t)16 * c);
    s.wft = reinterpret_cast<__nv_bfloat16 *>(p);
    return s;
}
template <int C>
void launch_fwd(long long rows, const float *pos, const PeState &s, const float *b2, float *out, int ga,
                float *aux_out, cudaStream_t st) {
    using L = PeTcLayout<C>;
    static bool once = (cudaFuncSetAttribute(pe_mlp_forward_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes), true);
    (void)once;
    pe_mlp_forward_tc_kernel<C><<<pe_grid(rows, kTcBlock, L::ctas_per_sm), kTcBlock, L::bytes, st>>>(
        rows, pos, s.fold, s.w2, b2, out, s.wf, ga, aux_out);
}
template <int C>
void launch_bwd(long long rows, const float *pos, const PeState &s, const float *grad, float *partial, int grid,
                int ga, const float *grad_aux, cudaStream_t st) {
    using L = PeTcBwd<C>;
    static bool once = (cudaFuncSetAttribute(pe_mlp_backward_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes), true);
    (void)once;
    pe_mlp_backward_tc_kernel<C><<<grid, kTcBlock, L::bytes, st>>>(rows, pos, s.fold, s.w2t, grad, partial, s.wft, ga, grad_aux);
}
}  // namespace

/* peb (rows,c) = W2·relu(BN(W1·pos + b1)) + b2.  use_batch_stats = 1 (training): BatchNorm statistics from
 * `moments` (aopt_pos_moments of the same pos); 0: running_mean / running_var.  stats_out (3c floats) receives
 * the batch mean, biased variance and rstd of the first layer (for the running-statistics update). */
extern "C" int aopt_pe_mlp_forward(int64_t rows, int c, const float *pos, const double *moments, const float *w1,
                                   const float *b1, const float *gamma, const float *beta,
                                   const float *running_mean, const float *running_var, float eps,
                                   int use_batch_stats, const float *w2, const float *b2, float *out,
                                   const float *aux_w, int ga, float *aux_out, void *state, size_t state_bytes,
                                   aopt_stream_t stream) {
    if (rows < 0 || !aopt_pe_mlp_supported(c)) return rows < 0 ? AOPT_ERR_INVALID_ARGUMENT : AOPT_ERR_UNSUPPORTED;
    if (!w1 || !b1 || !gamma || !beta || !w2 || !b2 || !state) return AOPT_ERR_INVALID_ARGUMENT;
    if (use_batch_stats ? !moments : (!running_mean || !running_var)) return AOPT_ERR_INVALID_ARGUMENT;
    if (state_bytes < aopt_pe_mlp_state_bytes(c)) return AOPT_ERR_WORKSPACE;
    if (rows > 0 && (!pos || !out)) return AOPT_ERR_INVALID_ARGUMENT;
    if (aux_w && (ga < 1 || ga > 16 || (rows > 0 && !aux_out))) return AOPT_ERR_INVALID_ARGUMENT;
    if (!aux_w) { ga = 0; aux_out = nullptr; }
    cudaStream_t st = as_stream(stream);
    PeState s = carve_state(state, c);
    pe_fold_kernel<<<c, 128, 0, st>>>(c, (double)rows, moments, w1, b1, gamma, beta, running_mean, running_var, eps,
                                      use_batch_stats, w2, s.fold, s.stats, s.w2, s.w2t, aux_w, ga, s.wf, s.wft);
    if (rows > 0) {
        if (c == 48) launch_fwd<48>(rows, pos, s, b2, out, ga, aux_out, st);
        else launch_fwd<96>(rows, pos, s, b2, out, ga, aux_out, st);
    }
    return check_launch(rows > 0 ? 2 : 1);
}

extern "C" size_t aopt_pe_mlp_backward_workspace_bytes(int64_t rows, int c) {
    if (!aopt_pe_mlp_supported(c) || rows < 0) return 0;
    const size_t pf = (size_t)c * c + 6 * (size_t)c + 16 * (size_t)c;
    return a256(4 * pf * (size_t)pe_bwd_grid(rows, c)) + a256(4 * 6 * (size_t)c);
}

/* Parameter gradients of aopt_pe_mlp_forward given grad (rows,c) = dL/dpeb; `state` is the forward's. */
extern "C" int aopt_pe_mlp_backward(int64_t rows, int c, const float *pos, const double *moments, const float *w1,
                                    const float *gamma, int use_batch_stats, const float *grad, const void *state,
                                    float *grad_w1, float *grad_b1, float *grad_gamma, float *grad_beta,
                                    float *grad_w2, float *grad_b2, int ga, const float *grad_aux,
                                    float *grad_aux_w, void *workspace, size_t workspace_bytes,
                                    aopt_stream_t stream) {
    if (rows < 1) return AOPT_ERR_INVALID_ARGUMENT;
    if (!aopt_pe_mlp_supported(c)) return AOPT_ERR_UNSUPPORTED;
    if (!pos || !w1 || !gamma || !grad || !state || !grad_w1 || !grad_b1 || !grad_gamma || !grad_beta || !grad_w2 || !grad_b2)
        return AOPT_ERR_INVALID_ARGUMENT;
    if (use_batch_stats && !moments) return AOPT_ERR_INVALID_ARGUMENT;
    if (grad_aux && (ga < 1 || ga > 16 || !grad_aux_w)) return AOPT_ERR_INVALID_ARGUMENT;
    if (!grad_aux) { ga = 0; grad_aux_w = nullptr; }
    if (!workspace || workspace_bytes < aopt_pe_mlp_backward_workspace_bytes(rows, c)) return AOPT_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    PeState s = carve_state(const_cast<void *>(state), c);
    const int grid = pe_bwd_grid(rows, c);
    const int pf = c * c + 6 * c + 16 * c;
    float *partial = static_cast<float *>(workspace);
    float *sums = reinterpret_cast<float *>(static_cast<char *>(workspace) + a256(4 * (size_t)pf * grid));
    if (c == 48) launch_bwd<48>(rows, pos, s, grad, partial, grid, ga, grad_aux, st);
    else launch_bwd<96>(rows, pos, s, grad, partial, grid, ga, grad_aux, st);
    pe_mlp_backward_finalize_kernel<<<div_up(pf, 32), 256, 0, st>>>(c, grid, pf, partial, grad_w2, sums, ga, grad_aux_w);
    pe_mlp_backward_params_kernel<<<div_up(c, 128), 128, 0, st>>>(c, (double)rows, use_batch_stats, sums, moments, w1,
                                                                  gamma, s.stats, grad_w1, grad_b1, grad_gamma,
                                                                  grad_beta, grad_b2);
    return check_launch(3);
}

/* stats (3c floats: mean | biased var | rstd) of the forward that filled `state`. */
extern "C" int aopt_pe_mlp_stats(int c, const void *state, float *stats_out, aopt_stream_t stream) {
    if (!aopt_pe_mlp_supported(c)) return AOPT_ERR_UNSUPPORTED;
    if (!state || !stats_out) return AOPT_ERR_INVALID_ARGUMENT;
    PeState s = carve_state(const_cast<void *>(state), c);
    cudaMemcpyAsync(stats_out, s.stats, 4 * (size_t)3 * c, cudaMemcpyDeviceToDevice, as_stream(stream));
    return check_launch(0);
}
