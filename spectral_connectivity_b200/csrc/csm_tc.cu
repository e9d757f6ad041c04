// Tensor-core cross-spectral matrix: C(b,f) = scale * sum_r X_r X_r^H on tcgen05 (sm_100a).
//
// Replaces the same reference code as csm.cu (connectivity.py:447-526, the k=1 batched matmul of
// :1799-1822 plus the xp.mean of :67-75) for SC_CSM_CROSS when S is large enough to fill a
// 128x128 UMMA tile.  Per (b, f) the planar coefficients are two real [R][S] slabs (Re, Im) with
// the signal axis contiguous, i.e. MN-major operands:
//
//     Re C = Ar Ar^T + Ai Ai^T          Im C = Ai Ar^T - Ar Ai^T
//
// so one complex tile is four real 128x128xR GEMMs on the same four operand slabs; the minus
// sign is the instruction descriptor's negate-A bit.  fp32 accuracy comes from the 3xTF32 split
// x = hi + lo (hi = tf32 truncation, lo = x - hi): hi*hi + hi*lo + lo*hi, the dropped lo*lo term
// is 2^-22 relative.  Only upper-triangular tiles are computed; the epilogue writes the tile and
// its conjugate transpose.
//
// Warp-specialised persistent kernel, one CTA per SM:
//   warp 0      TMA producer: 4-D tensor map {32 s, R, S/32, B*F*2}, swizzle 128B_ATOM_32B; a box
//               {32, KC, 4, 1} lands as [4][KC][32 f32] = the canonical MN-major tf32 layout of the
//               UMMA shared-memory descriptor; rows beyond R are zero-filled
//   warps 4-7   converter: split the landed fp32 slab into hi (in place) and lo (mirror buffer)
//   warp 1      MMA issuer: 12 tcgen05.mma.kind::tf32 (M=128, N=128, K=8) per K-step into a TMEM
//               accumulator pair (Re, Im); two pairs alternate per smem stage (4 x 128 = 512 columns)
//   warps 8-15  epilogue: drain every stage's partial tile with tcgen05.ld into fp32 registers
//               (the tensor core accumulates round-toward-zero), finally scale, store C and conj(C)^T
// Pipelines: smem ring (full_raw -> full_cvt -> empty) and TMEM ring (tmem_full/tmem_empty), all
// mbarriers.  Every spin is bounded and traps, so a protocol bug cannot hang the device.
//
// TWO kernels live here: csm_tc_kernel (both operands from shared memory, the round-1 design described above, kept behind
// SC_CSM_TA=0) and csm_tc_ta_kernel (DEFAULT: the row-block operand comes from tensor memory, see its own header below).
// Both issue their MMAs from a warp-uniform loop with the instruction predicated on the elected lane.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "sc_common.cuh"

namespace {

constexpr int TM = 128;                      // tile rows (signal i)
constexpr int TN = 128;                      // tile cols (signal j)
constexpr int KC = 16;                       // observations per smem stage
constexpr int STAGES = 3;
#ifndef SC_CSM_DRAIN
#define SC_CSM_DRAIN 1
#endif
constexpr int DRAIN = SC_CSM_DRAIN;         // smem stages accumulated in TMEM before the epilogue warps drain them
constexpr int SLAB = 4 * KC * 128;           // [4 groups of 32 signals][KC][128 B] = 8 KB
constexpr int STAGE_RAW = 4 * SLAB;          // Ar_I, Ai_I, Ar_J, Ai_J
constexpr int STAGE_BYTES = 2 * STAGE_RAW;   // + the lo mirror
constexpr int NTHREADS = 512;
constexpr int CVT_THREADS = 128;
constexpr int EPI_THREADS = 256;
constexpr uint32_t TMEM_COLS = 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (spin > (1u << 28)) __trap();  // watchdog: a broken pipeline must not hang the GPU
    }
}

#ifdef SC_CSM_PROFILE
#define PWAIT(acc, bar, par) do { const long long t0_ = clock64(); mbar_wait(bar, par); acc += clock64() - t0_; } while (0)
#else
#define PWAIT(acc, bar, par) mbar_wait(bar, par)  // acc only exists in profile builds
#endif
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// shared-memory matrix descriptor, MN-major 32-bit operands.  For tf32 the MN-major canonical layout is
// the "128B swizzle with 32B atoms" one (layout type 1, found by experiment: type 2 yields zeros):
// rows of 128 B (32 signals) whose four 32-byte chunks are XOR-permuted by (row & 3), K groups of 4 rows.
// Bit layout (PTX matrix descriptor): start address>>4 [0,14), leading byte offset>>4 [16,30) = stride
// between 32-element MN groups, stride byte offset>>4 [32,46) = stride between 4-row K groups,
// constant 0b001 [46,49), swizzle/layout type [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    constexpr uint64_t LBO = (KC * 128) >> 4;
    constexpr uint64_t SBO = 512 >> 4;
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (LBO << 16) | (SBO << 32) | (1ull << 46) | (1ull << 61);
}

// instruction descriptor: D=f32, A=B=tf32, both MN-major, M=128, N=128; optional negate-A
__host__ __device__ constexpr uint32_t umma_idesc(bool neg_a) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((neg_a ? 1u : 0u) << 13) | (1u << 15) | (1u << 16) |
           ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

// The MMA warp runs its loop with ALL 32 lanes (warp-uniform control flow and operands) and only the instruction itself
// is predicated on the elected lane: issued from inside `if (lane == 0)` every tcgen05.mma was wrapped by the compiler
// in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop that moves its operands into uniform registers (~10 instructions and a
// backward branch per MMA; measured with per-role cycle counters: the issuing thread was busy 2230 cycles per
// 24-MMA stage against 1536 cycles of tensor-pipe work, and the tensor pipe idled at 49 %).
__device__ __forceinline__ uint32_t elect_leader() {
    uint32_t is_leader;
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(is_leader));
    return is_leader;
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc,
                                          uint32_t leader) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(leader)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar, uint32_t leader) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar)), "r"(leader)
        : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

struct TcParams {
    long long BF, R, S;
    float scale;
    float2* out;
    int ntile;         // ceil(S / 128)
    long long ntiles;  // BF * ntile*(ntile+1)/2
};

__device__ __forceinline__ void decode_tile(long long t, int ntile, long long& bf, int& ti, int& tj) {
    const int npair = ntile * (ntile + 1) / 2;
    bf = t / npair;
    int pidx = (int)(t % npair);
    ti = 0;
    while (pidx >= ntile - ti) {
        pidx -= ntile - ti;
        ++ti;
    }
    tj = ti + pidx;
}

// Warp roles (16 warps = 4 warpgroups, registers rebalanced with setmaxnreg):
//   WG0: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2-3 idle      (48 regs)
//   WG1: warps 4-7 converter                                                       (48 regs)
//   WG2: warps 8-11 epilogue, columns 0-63 of the tile                             (208 regs)
//   WG3: warps 12-15 epilogue, columns 64-127                                      (208 regs)
// The tensor core accumulates with round-toward-zero (measured: -3.9e-8 relative bias per
// observation on coherent sums), so each smem stage (KC observations) goes to a fresh TMEM
// accumulator that the epilogue warps drain into fp32 registers with round-to-nearest adds; within
// a stage the small hi*lo / lo*hi products are issued first, while the accumulator is still small.
__global__ void __launch_bounds__(NTHREADS, 1) csm_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TcParams p) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full_raw[STAGES], full_cvt[STAGES], empty_bar[STAGES], tmem_full[2], tmem_empty[2];
    __shared__ uint32_t tmem_base_sh;

    unsigned char* stage_base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = (int)((p.R + KC - 1) / KC);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_raw[s], 1);
            mbar_init(&full_cvt[s], CVT_THREADS);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], EPI_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_sh;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    }
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
                long long bf;
                int ti, tj;
                decode_tile(t, p.ntile, bf, ti, tj);
                const bool diag = ti == tj;
                for (int kc = 0; kc < nk; ++kc) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    unsigned char* dst = stage_base + (size_t)stage * STAGE_BYTES;
                    mbar_expect_tx(&full_raw[stage], (diag ? 2 : 4) * SLAB);
                    const int plane = (int)(bf * 2);
                    tma_load_4d(dst, &tmap, &full_raw[stage], 0, kc * KC, ti * 4, plane);
                    tma_load_4d(dst + SLAB, &tmap, &full_raw[stage], 0, kc * KC, ti * 4, plane + 1);
                    if (!diag) {
                        tma_load_4d(dst + 2 * SLAB, &tmap, &full_raw[stage], 0, kc * KC, tj * 4, plane);
                        tma_load_4d(dst + 3 * SLAB, &tmap, &full_raw[stage], 0, kc * KC, tj * 4, plane + 1);
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp, elected lane issues) =====================
        {
            const uint32_t leader = elect_leader();
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            constexpr uint32_t idesc = umma_idesc(false), idesc_neg = umma_idesc(true);
            for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
                long long bf;
                int ti, tj;
                decode_tile(t, p.ntile, bf, ti, tj);
                const bool diag = ti == tj;
                for (int kc = 0; kc < nk; ++kc) {
                    const bool fresh = kc % DRAIN == 0;                       // first stage of an accumulator
                    const bool last = kc % DRAIN == DRAIN - 1 || kc == nk - 1;  // ... and its last
                    if (fresh) mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                    mbar_wait(&full_cvt[stage], phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t d_re = tmem_base + (uint32_t)acc * 256u;
                    const uint32_t d_im = d_re + 128u;
                    const uint32_t hi = smem_u32(stage_base + (size_t)stage * STAGE_BYTES);
                    const uint32_t lo = hi + STAGE_RAW;
                    const uint32_t boff = diag ? 0u : 2u * SLAB;
                    // pass 0: cross terms hi*lo + lo*hi (small); pass 1: hi*hi (large)
#pragma unroll
                    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
                        for (int ks = 0; ks < KC / 8; ++ks) {
                            const uint32_t ko = (uint32_t)ks * 1024u;
                            const uint64_t arh = umma_desc(hi + ko), aih = umma_desc(hi + SLAB + ko);
                            const uint64_t brh = umma_desc(hi + boff + ko), bih = umma_desc(hi + boff + SLAB + ko);
                            if (pass == 0) {
                                const uint64_t arl = umma_desc(lo + ko), ail = umma_desc(lo + SLAB + ko);
                                const uint64_t brl = umma_desc(lo + boff + ko), bil = umma_desc(lo + boff + SLAB + ko);
                                const uint32_t first = (ks || !fresh) ? 1u : 0u;  // fresh accumulator every DRAIN stages
                                umma_tf32(d_re, arh, brl, idesc, first, leader);
                                umma_tf32(d_re, arl, brh, idesc, 1u, leader);
                                umma_tf32(d_re, aih, bil, idesc, 1u, leader);
                                umma_tf32(d_re, ail, bih, idesc, 1u, leader);
                                umma_tf32(d_im, aih, brl, idesc, first, leader);
                                umma_tf32(d_im, ail, brh, idesc, 1u, leader);
                                umma_tf32(d_im, arh, bil, idesc_neg, 1u, leader);
                                umma_tf32(d_im, arl, bih, idesc_neg, 1u, leader);
                            } else {
                                // Re += Ar Br^T + Ai Bi^T ;  Im += Ai Br^T - Ar Bi^T
                                umma_tf32(d_re, arh, brh, idesc, 1u, leader);
                                umma_tf32(d_re, aih, bih, idesc, 1u, leader);
                                umma_tf32(d_im, aih, brh, idesc, 1u, leader);
                                umma_tf32(d_im, arh, bih, idesc_neg, 1u, leader);
                            }
                        }
                    }
                    umma_commit(&empty_bar[stage], leader);  // smem stage reusable once these MMAs retire
                    if (last) umma_commit(&tmem_full[acc], leader);  // and this partial tile can be drained
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                    if (last && ++acc == 2) {
                        acc = 0;
                        acc_phase ^= 1;
                    }
                }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===================== converter: fp32 -> (hi, lo) =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
        const int ct = threadIdx.x - 128;
        int stage = 0;
        uint32_t phase = 0;
        for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            long long bf;
            int ti, tj;
            decode_tile(t, p.ntile, bf, ti, tj);
            const int nvec = (ti == tj ? 2 : 4) * SLAB / 16;
            for (int kc = 0; kc < nk; ++kc) {
                mbar_wait(&full_raw[stage], phase);
                float4* raw = reinterpret_cast<float4*>(stage_base + (size_t)stage * STAGE_BYTES);
                float4* lo = reinterpret_cast<float4*>(stage_base + (size_t)stage * STAGE_BYTES + STAGE_RAW);
#pragma unroll 4
                for (int idx = ct; idx < nvec; idx += CVT_THREADS) {
                    const float4 v = raw[idx];
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
                    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
                    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
#ifndef SC_CSM_KEEP_RAW_HI
                    raw[idx] = h;
#endif
                    lo[idx] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&full_cvt[stage]);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp >= 8) {
        // ===================== epilogue / drain =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        const int q = warp & 3;           // TMEM lane quarter this warp may access
        const int half = (warp - 8) >> 2;  // column half of the tile
        int acc = 0;
        uint32_t acc_phase = 0;
        for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            long long bf;
            int ti, tj;
            decode_tile(t, p.ntile, bf, ti, tj);
            const bool diag = ti == tj;
            float sre[64], sim[64];
#pragma unroll
            for (int u = 0; u < 64; ++u) {
                sre[u] = 0.f;
                sim[u] = 0.f;
            }
            for (int kc = 0; kc < nk; kc += DRAIN) {
                mbar_wait(&tmem_full[acc], acc_phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u + (uint32_t)half * 64u;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t re[32], im[32];
                    tmem_ld32(trow + c * 32, re);
                    tmem_ld32(trow + 128 + c * 32, im);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (c == 1) {
                        // this accumulator has been read completely: hand it back to the MMA warp
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        mbar_arrive(&tmem_empty[acc]);
                    }
#pragma unroll
                    for (int u = 0; u < 32; ++u) {
                        sre[c * 32 + u] += __uint_as_float(re[u]);
                        sim[c * 32 + u] += __uint_as_float(im[u]);
                    }
                }
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
            const long long i = (long long)ti * TM + q * 32 + lane;
            const long long j0 = (long long)tj * TN + half * 64;
            float2* mat = p.out + bf * p.S * p.S;
#pragma unroll
            for (int u = 0; u < 64; ++u) {
                const long long j = j0 + u;
                const float2 v = make_float2(sre[u] * p.scale, sim[u] * p.scale);
                if (j < p.S && i < p.S) {
                    mat[i * p.S + j] = v;
                    if (!diag) mat[j * p.S + i] = make_float2(v.x, -v.y);  // coalesced across lanes
                }
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}


// =====================================================================================================================
// Variant with the A operand in TENSOR MEMORY (default; SC_CSM_TA=0 selects the kernel above).
//
// A 128x128x8 tf32 MMA with both operands in shared memory reads 4 KB (A) + 4 KB (B) per 64 clocks = 128 B/clk, the whole
// shared-memory bandwidth of the SM, so the kernel above can never exceed the share of that bandwidth the TMA writes and
// the converter warps leave it (measured: tensor pipe 49 %; draining the accumulators 4x less often or dropping the
// hi write-back moved it by 8 % / 2 %).  Here the converter warps write the hi/lo split of the ROW block straight into
// TMEM (tcgen05.st, lane = signal row, 8 columns per K step and operand) and the MMAs take A from there
// (tcgen05.mma [d], [a], b-desc): operand reads from shared memory halve (96 KB instead of 192 KB per 16-observation
// stage), the row block needs no lo mirror in shared memory, and the hi operand of the column block is the RAW fp32
// slab (kind::tf32 ignores the low 13 mantissa bits -- measured: results identical to an explicit truncation).
// TMEM: three 128-column accumulators used as a ring of HALF stages (Re, Im, Re, ...) + two 64-column A buffers.
// Shared memory: 4 stages of {raw I (re, im), raw J (re, im), lo J (re, im)} = 48 KB.
#ifndef SC_CSM_TA_STAGES
#define SC_CSM_TA_STAGES 4
#endif
constexpr int TA_STAGES = SC_CSM_TA_STAGES;
constexpr int TA_STAGE_BYTES = 6 * SLAB;
constexpr uint32_t TA_ACC_SLOTS = 3;
constexpr uint32_t TA_A_BASE = 384;  // TMEM column of the first A buffer (2 x 64 columns)

#ifdef SC_CSM_FAKE_KMAJOR  // TIMING EXPERIMENT ONLY (wrong results): B declared K-major, 64-byte swizzle
__host__ __device__ constexpr uint32_t umma_idesc_ta(bool neg_a) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((neg_a ? 1u : 0u) << 13) | ((uint32_t)(TN >> 3) << 17) |
           ((uint32_t)(TM >> 4) << 24);
}
__device__ __forceinline__ uint64_t umma_desc_b(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
#else
__host__ __device__ constexpr uint32_t umma_idesc_ta(bool neg_a) {  // A from TMEM (K-major), B MN-major
    return (1u << 4) | (2u << 7) | (2u << 10) | ((neg_a ? 1u : 0u) << 13) | (1u << 16) | ((uint32_t)(TN >> 3) << 17) |
           ((uint32_t)(TM >> 4) << 24);
}
__device__ __forceinline__ uint64_t umma_desc_b(uint32_t saddr) { return umma_desc(saddr); }
#endif
#ifdef SC_CSM_FAKE_KMAJOR
constexpr uint32_t TA_KSTEP_BYTES = 32u;
#else
constexpr uint32_t TA_KSTEP_BYTES = 1024u;  // 8 observations x 128-byte rows
#endif

__device__ __forceinline__ void umma_tf32_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc,
                                             uint32_t leader) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc), "r"(leader)
        : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1) csm_tc_ta_kernel(const __grid_constant__ CUtensorMap tmap, const TcParams p) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full_raw[TA_STAGES], full_cvt[TA_STAGES], empty_bar[TA_STAGES], a_empty[2],
        acc_full[TA_ACC_SLOTS], acc_empty[TA_ACC_SLOTS];
    __shared__ uint32_t tmem_base_sh;

    unsigned char* stage_base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = (int)((p.R + KC - 1) / KC);
#ifdef SC_CSM_PROFILE  // per-role mbarrier wait accounting (tools/csm_role_profile.py)
    long long w0 = 0, w1 = 0, w2 = 0;
    const long long tstart = clock64();
#endif

    if (threadIdx.x == 0) {
        for (int s = 0; s < TA_STAGES; ++s) {
            mbar_init(&full_raw[s], 1);
            mbar_init(&full_cvt[s], CVT_THREADS);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) mbar_init(&a_empty[a], 1);
        for (int a = 0; a < (int)TA_ACC_SLOTS; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], EPI_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_sh;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    }
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
                long long bf;
                int ti, tj;
                decode_tile(t, p.ntile, bf, ti, tj);
                const bool diag = ti == tj;
                for (int kc = 0; kc < nk; ++kc) {
                    PWAIT(w0, &empty_bar[stage], phase ^ 1);
                    unsigned char* dst = stage_base + (size_t)stage * TA_STAGE_BYTES;
                    mbar_expect_tx(&full_raw[stage], (diag ? 2 : 4) * SLAB);
                    const int plane = (int)(bf * 2);
                    tma_load_4d(dst + 2 * SLAB, &tmap, &full_raw[stage], 0, kc * KC, tj * 4, plane);
                    tma_load_4d(dst + 3 * SLAB, &tmap, &full_raw[stage], 0, kc * KC, tj * 4, plane + 1);
                    if (!diag) {
                        tma_load_4d(dst, &tmap, &full_raw[stage], 0, kc * KC, ti * 4, plane);
                        tma_load_4d(dst + SLAB, &tmap, &full_raw[stage], 0, kc * KC, ti * 4, plane + 1);
                    }
                    if (++stage == TA_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp, elected lane issues) =====================
        {
            const uint32_t leader = elect_leader();
            int stage = 0;
            uint32_t phase = 0;
            uint32_t slot = 0, slot_phase = 0;   // accumulator ring (half stages)
            uint32_t ab = 0;                     // A buffer of this stage
            constexpr uint32_t idesc = umma_idesc_ta(false), idesc_neg = umma_idesc_ta(true);
            for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
                for (int kc = 0; kc < nk; ++kc) {
                    PWAIT(w0, &full_cvt[stage], phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sb = smem_u32(stage_base + (size_t)stage * TA_STAGE_BYTES);
                    const uint32_t brh_s = sb + 2u * SLAB, bih_s = sb + 3u * SLAB, brl_s = sb + 4u * SLAB, bil_s = sb + 5u * SLAB;
                    const uint32_t a0 = tmem_base + TA_A_BASE + ab * 64u;  // [ks][Ar_hi, Ar_lo, Ai_hi, Ai_lo][8 columns]
                    // ---- Re = Ar Br^T + Ai Bi^T: cross terms first (small), then hi*hi ----
                    PWAIT(w1, &acc_empty[slot], slot_phase ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    {
                        const uint32_t d = tmem_base + slot * 128u;
#pragma unroll
                        for (int ks = 0; ks < KC / 8; ++ks) {
                            const uint32_t ko = (uint32_t)ks * TA_KSTEP_BYTES, ak = a0 + (uint32_t)ks * 32u;
                            umma_tf32_ta(d, ak + 0u, umma_desc_b(brl_s + ko), idesc, ks ? 1u : 0u, leader);
                            umma_tf32_ta(d, ak + 8u, umma_desc_b(brh_s + ko), idesc, 1u, leader);
                            umma_tf32_ta(d, ak + 16u, umma_desc_b(bil_s + ko), idesc, 1u, leader);
                            umma_tf32_ta(d, ak + 24u, umma_desc_b(bih_s + ko), idesc, 1u, leader);
                        }
#pragma unroll
                        for (int ks = 0; ks < KC / 8; ++ks) {
                            const uint32_t ko = (uint32_t)ks * TA_KSTEP_BYTES, ak = a0 + (uint32_t)ks * 32u;
                            umma_tf32_ta(d, ak + 0u, umma_desc_b(brh_s + ko), idesc, 1u, leader);
                            umma_tf32_ta(d, ak + 16u, umma_desc_b(bih_s + ko), idesc, 1u, leader);
                        }
                        umma_commit(&acc_full[slot], leader);
                        if (++slot == TA_ACC_SLOTS) {
                            slot = 0;
                            slot_phase ^= 1;
                        }
                    }
                    // ---- Im = Ai Br^T - Ar Bi^T ----
                    PWAIT(w1, &acc_empty[slot], slot_phase ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    {
                        const uint32_t d = tmem_base + slot * 128u;
#pragma unroll
                        for (int ks = 0; ks < KC / 8; ++ks) {
                            const uint32_t ko = (uint32_t)ks * TA_KSTEP_BYTES, ak = a0 + (uint32_t)ks * 32u;
                            umma_tf32_ta(d, ak + 16u, umma_desc_b(brl_s + ko), idesc, ks ? 1u : 0u, leader);
                            umma_tf32_ta(d, ak + 24u, umma_desc_b(brh_s + ko), idesc, 1u, leader);
                            umma_tf32_ta(d, ak + 0u, umma_desc_b(bil_s + ko), idesc_neg, 1u, leader);
                            umma_tf32_ta(d, ak + 8u, umma_desc_b(bih_s + ko), idesc_neg, 1u, leader);
                        }
#pragma unroll
                        for (int ks = 0; ks < KC / 8; ++ks) {
                            const uint32_t ko = (uint32_t)ks * TA_KSTEP_BYTES, ak = a0 + (uint32_t)ks * 32u;
                            umma_tf32_ta(d, ak + 16u, umma_desc_b(brh_s + ko), idesc, 1u, leader);
                            umma_tf32_ta(d, ak + 0u, umma_desc_b(bih_s + ko), idesc_neg, 1u, leader);
                        }
                        umma_commit(&acc_full[slot], leader);
                        if (++slot == TA_ACC_SLOTS) {
                            slot = 0;
                            slot_phase ^= 1;
                        }
                    }
                    umma_commit(&empty_bar[stage], leader);  // shared-memory stage and A buffer reusable once these MMAs retire
                    umma_commit(&a_empty[ab], leader);
                    ab ^= 1u;
                    if (++stage == TA_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===================== converter: row block -> TMEM (hi, lo); column block -> lo mirror =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
        const int ct = threadIdx.x - 128;            // signal row of the tile = TMEM lane
        const int g = ct >> 5, sgn = ct & 31;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        int stage = 0;
        uint32_t phase = 0, ab = 0, ab_phase = 0;
        for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            long long bf;
            int ti, tj;
            decode_tile(t, p.ntile, bf, ti, tj);
            const bool diag = ti == tj;
            for (int kc = 0; kc < nk; ++kc) {
                PWAIT(w0, &full_raw[stage], phase);
                unsigned char* sbase = stage_base + (size_t)stage * TA_STAGE_BYTES;
                const unsigned char* src = sbase + (diag ? 2 * SLAB : 0);  // raw row block (re, then im)
                PWAIT(w1, &a_empty[ab], ab_phase ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = tmem_base + lane_base + TA_A_BASE + ab * 64u;
#pragma unroll
                for (int ks = 0; ks < KC / 8; ++ks) {
#pragma unroll
                    for (int pl = 0; pl < 2; ++pl) {
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = ks * 8 + i;
                            // TMA box {32 s, KC r, 4 groups}, swizzle 128B with 32-byte atoms: chunk ^= (r & 3)
                            const float x = *reinterpret_cast<const float*>(
                                src + pl * SLAB + g * (KC * 128) + r * 128 + ((((sgn >> 3) ^ (r & 3))) << 5) + (sgn & 7) * 4);
                            const uint32_t h = __float_as_uint(x) & 0xffffe000u;
                            hi[i] = h;
                            lo[i] = __float_as_uint(x - __uint_as_float(h));
                        }
                        tmem_st8(a0 + (uint32_t)(ks * 32 + pl * 16), hi);
                        tmem_st8(a0 + (uint32_t)(ks * 32 + pl * 16 + 8), lo);
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                // lo mirror of the column block (its hi operand is the raw slab itself)
                const float4* raw = reinterpret_cast<const float4*>(sbase + 2 * SLAB);
                float4* lom = reinterpret_cast<float4*>(sbase + 4 * SLAB);
#pragma unroll 4
                for (int idx = ct; idx < 2 * SLAB / 16; idx += CVT_THREADS) {
                    const float4 v = raw[idx];
                    float4 l;
                    l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
                    l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                    l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
                    l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                    lom[idx] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(&full_cvt[stage]);
                ab ^= 1u;
                if (ab == 0) ab_phase ^= 1;
                if (++stage == TA_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp >= 8) {
        // ===================== epilogue / drain =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        const int q = warp & 3;            // TMEM lane quarter this warp may access
        const int half = (warp - 8) >> 2;  // column half of the tile
        uint32_t slot = 0, slot_phase = 0;
        for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            long long bf;
            int ti, tj;
            decode_tile(t, p.ntile, bf, ti, tj);
            const bool diag = ti == tj;
            float sre[64], sim[64];
#pragma unroll
            for (int u = 0; u < 64; ++u) {
                sre[u] = 0.f;
                sim[u] = 0.f;
            }
            for (int kc = 0; kc < nk; ++kc) {
#pragma unroll
                for (int part = 0; part < 2; ++part) {  // 0: Re half stage, 1: Im half stage
                    PWAIT(w0, &acc_full[slot], slot_phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + slot * 128u + (uint32_t)half * 64u;
                    uint32_t v0[32], v1[32];
                    tmem_ld32(trow, v0);
                    tmem_ld32(trow + 32, v1);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(&acc_empty[slot]);
                    if (part == 0) {
#pragma unroll
                        for (int u = 0; u < 32; ++u) {
                            sre[u] += __uint_as_float(v0[u]);
                            sre[32 + u] += __uint_as_float(v1[u]);
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < 32; ++u) {
                            sim[u] += __uint_as_float(v0[u]);
                            sim[32 + u] += __uint_as_float(v1[u]);
                        }
                    }
                    if (++slot == TA_ACC_SLOTS) {
                        slot = 0;
                        slot_phase ^= 1;
                    }
                }
            }
            const long long i = (long long)ti * TM + q * 32 + lane;
            const long long j0 = (long long)tj * TN + half * 64;
            float2* mat = p.out + bf * p.S * p.S;
#pragma unroll
            for (int u = 0; u < 64; ++u) {
                const long long j = j0 + u;
                const float2 v = make_float2(sre[u] * p.scale, sim[u] * p.scale);
                if (j < p.S && i < p.S) {
                    mat[i * p.S + j] = v;
                    if (!diag) mat[j * p.S + i] = make_float2(v.x, -v.y);  // coalesced across lanes
                }
            }
        }
    }

#ifdef SC_CSM_PROFILE
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 1 || warp == 4 || warp == 8))
        printf("warp %d total %lld wait0 %lld wait1 %lld wait2 %lld stages %lld\n", warp, clock64() - tstart, w0, w1, w2,
               (long long)nk * ((p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x));
#endif
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(sym);
    return fn;
}

}  // namespace

int sc_csm_tc_supported(int64_t R, int64_t S) {
    static int disabled = -1;
    if (disabled < 0) {
        const char* e = getenv("SC_B200_DISABLE_TC");
        disabled = (e && e[0] == '1') ? 1 : 0;
    }
    if (disabled) return 0;
    // S must tile into 32-signal TMA groups; below 96 signals a 128x128 UMMA tile is mostly padding
    return S >= 96 && S % 32 == 0 && R >= 1 && get_encode() != nullptr;
}

int sc_csm_tc_launch(const float* xp, int64_t B, int64_t F, int64_t R, int64_t S, float scale, void* out,
                     cudaStream_t st) {
    SC_CHECK_ARG(xp && out, "sc_csm[tc]: null pointer");
    SC_CHECK_ARG(B > 0 && F > 0 && R > 0 && S > 0 && S % 32 == 0, "sc_csm[tc]: unsupported shape");
    SC_CHECK_ARG((reinterpret_cast<uintptr_t>(xp) & 15) == 0, "sc_csm[tc]: coefficients must be 16-byte aligned");
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        sc_set_error("sc_csm[tc]: cuTensorMapEncodeTiled unavailable");
        return SC_ERR_UNSUPPORTED;
    }
    const long long planes = B * F * 2;
    SC_CHECK_ARG(planes < (1LL << 31) && R < (1LL << 31), "sc_csm[tc]: batch too large for one tensor map");
    CUtensorMap tmap;
    const cuuint64_t gdim[4] = {32, (cuuint64_t)R, (cuuint64_t)(S / 32), (cuuint64_t)planes};
    const cuuint64_t gstr[3] = {(cuuint64_t)S * 4, 128, (cuuint64_t)R * S * 4};  // bytes, dims 1..3
    const cuuint32_t box[4] = {32, KC, 4, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(xp), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        sc_set_error("sc_csm[tc]: cuTensorMapEncodeTiled failed with %d (R=%lld S=%lld planes=%lld)", (int)cr,
                     (long long)R, (long long)S, planes);
        return SC_ERR_CUDA;
    }
    TcParams p;
    p.BF = B * F; p.R = R; p.S = S; p.scale = scale; p.out = reinterpret_cast<float2*>(out);
    p.ntile = (int)((S + TM - 1) / TM);
    p.ntiles = p.BF * (long long)(p.ntile * (p.ntile + 1) / 2);
    long long grid = sc_num_sms();
    if (grid > p.ntiles) grid = p.ntiles;
    const char* ta = getenv("SC_CSM_TA");  // "0": both operands from shared memory (the first kernel)
    if (ta && ta[0] == '0') {
        const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024;
        SC_CUDA_OK(cudaFuncSetAttribute(csm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        csm_tc_kernel<<<(unsigned)grid, NTHREADS, smem, st>>>(tmap, p);
    } else {
        const size_t smem = (size_t)TA_STAGES * TA_STAGE_BYTES + 1024;
        SC_CUDA_OK(cudaFuncSetAttribute(csm_tc_ta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        csm_tc_ta_kernel<<<(unsigned)grid, NTHREADS, smem, st>>>(tmap, p);
    }
    SC_LAUNCH_OK();
    return SC_OK;
}
