/*
 * rx_ws.cuh -- warp-specialised fast decimating RX (8 samples per symbol, compile-time tap
 * count NT).  Replaces demodulator.rs:44-55 + the two fir.rs:18-34 filters of the reference
 * (paths relative to /root/reference/src/modem/), plus the decimator / slicer / error-count
 * extension.  Same arithmetic as every other RX kernel: bit-identical results.
 *
 * Why this shape.  With 2 symbols per thread and a carrier table in shared memory the L1 /
 * shared-memory data pipe ran at 86 % (ncu, profiles/): FIR loads 576 + carrier loads 131 +
 * staging stores 131 + global loads ~290 wavefronts per 2048-sample tile.  Here
 *   - PRODUCER warps own a fixed slice of the tile (the same sample indices for every frame),
 *     so their (cos, -sin) NCO values live in REGISTERS for the whole frame loop: no carrier
 *     table, no carrier loads.  They stream frame i+1 from HBM (128-bit loads, all issued
 *     before the first use), multiply the real parts and store (vi, vq) pairs into the other
 *     half of a double-buffered shared tile while
 *   - CONSUMER warps run the FIR of frame i: every thread owns R = 4 consecutive symbols and
 *     walks its NB+R-1 eight-sample blocks once, newest first, so each 128-bit shared load
 *     feeds four symbols (22 B/sample of FIR traffic instead of 36); a MAC on both rails is
 *     FMUL2 + FFMA2 (common.cuh, f32x2), taps are uniform-register operands.
 *   - The two sides meet on named barriers (bar.sync / bar.arrive), not __syncthreads, so
 *     neither waits for the other's memory latency.
 *
 * Tile geometry (shared with rx_fast.cuh): block B holds tile-local samples [8B, 8B+7]; the
 * decision instant of tile symbol r is element 7-OFF of block r+NB-1 (OFF = 0 for an odd
 * decision delay, 1 for an even one, so blocks coincide with the 16-byte sample pairs), and
 * tap i of symbol r reads element e of block r+NB-1-b with i = 8b + 7 - OFF - e.
 * Shared layout: 16-byte chunk c (samples 2c, 2c+1) sits at chunk position c + c/(4R).
 */
#pragma once

#include "common.cuh"
#include "rx_fast.cuh" /* slice_point4 */

namespace mg {

template <int NT, int OFF, int R, int CW, int PW>
struct RxWsCfg {
    static constexpr int NB = (NT + OFF + 7) / 8;
    static constexpr int CT = 32 * CW;            /* consumer threads */
    static constexpr int PT = 32 * PW;            /* producer threads */
    static constexpr int THREADS = CT + PT;
    static constexpr int TS = R * CT;             /* symbols per tile */
    static constexpr int NBLK = TS + NB - 1;
    static constexpr int NSAMP = NBLK * 8;
    static constexpr int NCHUNK = NSAMP / 2;
    static constexpr int PADW = 4 * R;
    static constexpr int PCHUNK = NCHUNK + NCHUNK / PADW + 1;
    static constexpr int ITER = (NCHUNK + PT - 1) / PT; /* chunks per producer thread */
    static constexpr int NSTEP = NB + R - 1;
    static size_t smem(uint32_t lut_entries) { return 2 * sizeof(float4) * PCHUNK + sizeof(float2) * lut_entries; }
    static_assert(PT % PADW == 0, "producer position arithmetic needs PT % (4R) == 0");
};

/* named barriers with compile-time ids (a runtime id makes ptxas reserve all 16) */
template <int ID, int COUNT>
__device__ __forceinline__ void named_bar_sync()
{
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}
template <int ID, int COUNT>
__device__ __forceinline__ void named_bar_arrive()
{
    asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}
template <int ID0, int COUNT>
__device__ __forceinline__ void named_bar_sync2(int b)
{
    if (b) named_bar_sync<ID0 + 1, COUNT>();
    else named_bar_sync<ID0, COUNT>();
}
template <int ID0, int COUNT>
__device__ __forceinline__ void named_bar_arrive2(int b)
{
    if (b) named_bar_arrive<ID0 + 1, COUNT>();
    else named_bar_arrive<ID0, COUNT>();
}

template <int NT, int OFF, bool FMA, bool NOISE, int R, int CW, int PW, int MINB>
__global__ void __launch_bounds__(32 * (CW + PW), MINB)
    rx_ws_kernel(const __grid_constant__ RxArgs a, const __grid_constant__ TapsParam<NT> taps)
{
    using C = RxWsCfg<NT, OFF, R, CW, PW>;
    constexpr int BAR_FULL = 1, BAR_EMPTY = 3; /* + buffer index; barrier 0 is __syncthreads */
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* s_v = reinterpret_cast<float4*>(smem_raw); /* [2][PCHUNK] chunks of (vi0,vq0,vi1,vq1) */
    float2* s_slut = reinterpret_cast<float2*>(s_v + 2 * C::PCHUNK);

    const int tid = threadIdx.x;
    for (uint32_t i = tid; i < a.n_tables * a.n_const; i += C::THREADS) s_slut[i] = a.slut[i];
    __syncthreads();

    const u64 f0 = (u64)blockIdx.y * a.frames_per_block;
    const u64 f1 = min(a.F, f0 + a.frames_per_block);
    const int nframes = (int)(f1 - f0);
    const u64 k0 = (u64)blockIdx.x * C::TS;
    /* sample index of tile-local j = 0; even by the choice of OFF */
    const long long nbase = (long long)(k0 * 8 + a.delay) + OFF - 8 * C::NB + 1;
    uint32_t err = 0, cmp = 0;

    if (tid >= C::CT) {
        /* ============================ PRODUCER ============================ */
        const int ptid = tid - C::CT;
        const float w = chan_w(a.ch, f0), po = chan_po(a.ch, f0);
        const bool interior = nbase >= 0 && (u64)(nbase + C::NSAMP) <= a.L;
        float cs[C::ITER][4]; /* (c0, -s0, c1, -s1) of this thread's chunks: frame-invariant */
#pragma unroll
        for (int it = 0; it < C::ITER; ++it) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const long long n = nbase + 2 * (it * C::PT + ptid) + e;
                float s = 0.0f, c = 0.0f;
                if (it * C::PT + ptid < C::NCHUNK && n >= 0 && (u64)n < a.L)
                    mg_sincosf(__fadd_rn(nco_phase(w, a.sample0 + (u64)n), po), &s, &c);
                cs[it][2 * e] = c;
                cs[it][2 * e + 1] = -s;
            }
        }
        const int wbase = ptid + ptid / C::PADW; /* chunk position of chunk `ptid` */
        const float2* frame = a.rx + f0 * a.L;
        for (int i = 0; i < nframes; ++i, frame += a.L) {
            const int b = i & 1;
            float xr[C::ITER][2];
            if (interior) {
                const float4* src = reinterpret_cast<const float4*>(frame + nbase) + ptid;
#pragma unroll
                for (int it = 0; it < C::ITER; ++it) {
                    if (it * C::PT + ptid < C::NCHUNK) {
                        const float4 t = __ldcs(src + it * C::PT);
                        xr[it][0] = t.x;
                        xr[it][1] = t.z;
                    }
                }
            } else {
#pragma unroll
                for (int it = 0; it < C::ITER; ++it) {
                    const long long n = nbase + 2 * (it * C::PT + ptid);
                    xr[it][0] = 0.0f;
                    xr[it][1] = 0.0f;
                    if (it * C::PT + ptid < C::NCHUNK) {
                        if (n >= 0 && (u64)n < a.L) xr[it][0] = __ldcs(&frame[n].x);
                        if (n + 1 >= 0 && (u64)(n + 1) < a.L) xr[it][1] = __ldcs(&frame[n + 1].x);
                    }
                }
            }
            if (NOISE) {
                const u64 gf = a.nz.frame0 + f0 + i;
#pragma unroll 1
                for (int it = 0; it < C::ITER; ++it) {
                    const long long n = nbase + 2 * (it * C::PT + ptid);
                    if (it * C::PT + ptid < C::NCHUNK) {
                        float n0 = 0.0f, n1 = 0.0f;
                        const bool v0 = n >= 0 && (u64)n < a.L, v1 = n + 1 >= 0 && (u64)(n + 1) < a.L;
                        if (v0) n0 = noise_re(a.nz, gf, (u64)n);
                        if (v1) n1 = noise_re(a.nz, gf, (u64)(n + 1));
#pragma unroll
                        for (int j = 0; j < C::ITER; ++j) /* static index keeps xr[] in registers */
                            if (j == it) {
                                if (v0) xr[j][0] = __fadd_rn(xr[j][0], __fmul_rn(a.nz.sigma, n0));
                                if (v1) xr[j][1] = __fadd_rn(xr[j][1], __fmul_rn(a.nz.sigma, n1));
                            }
                    }
                }
            }
            /* the loads above are in flight while we wait for the consumers to release buffer b */
            if (i >= 2) named_bar_sync2<BAR_EMPTY, C::THREADS>(b);
            float4* dst = s_v + b * C::PCHUNK + wbase;
#pragma unroll
            for (int it = 0; it < C::ITER; ++it) {
                if (it * C::PT + ptid < C::NCHUNK)
                    /* chunk q = it*PT + ptid  ->  position q + q/PADW = wbase + it*(PT + PT/PADW) */
                    dst[it * (C::PT + C::PT / C::PADW)] =
                        make_float4(__fmul_rn(xr[it][0], cs[it][0]), __fmul_rn(xr[it][0], cs[it][1]),
                                    __fmul_rn(xr[it][1], cs[it][2]), __fmul_rn(xr[it][1], cs[it][3]));
            }
            named_bar_arrive2<BAR_FULL, C::THREADS>(b);
        }
    } else {
        /* ============================ CONSUMER ============================ */
        const f32x2 one = pk2(taps.one.x, taps.one.y);
        const u64 ka = k0 + (u64)R * tid; /* this thread's symbols: ka .. ka+R-1 */
        uint32_t toff[R];
#pragma unroll
        for (int r = 0; r < R; ++r) toff[r] = (uint32_t)((ka + r) % a.n_tables) * a.n_const;
        const bool have_any = ka < a.K, have_all = ka + R <= a.K;
        /* vector emit: every frame's symbol row must start on a multiple of R symbols */
        const bool vec_out = a.bps == 2 && have_all && (a.K % R == 0) &&
                             ((reinterpret_cast<uintptr_t>(a.sym) % R) == 0) &&
                             ((reinterpret_cast<uintptr_t>(a.bits) % (2 * R)) == 0);
        const bool ref_vec = a.ref_bits && vec_out && (a.ref_stride % (2 * R) == 0) &&
                             ((reinterpret_cast<uintptr_t>(a.ref_bits) % (2 * R)) == 0);
        u64 orow = f0 * a.K + ka;
        const uint8_t* refp = a.ref_bits ? a.ref_bits + f0 * a.ref_stride + ka * 2 : nullptr;
        for (int i = 0; i < nframes; ++i, orow += a.K) {
            const int b = i & 1;
            uint2 refw = make_uint2(0u, 0u);
            if (ref_vec) { /* issued before the wait: it arrives while the FIR runs */
                if (R == 4) refw = __ldg(reinterpret_cast<const uint2*>(refp));
                else refw.x = __ldg(reinterpret_cast<const uint32_t*>(refp));
            }
            named_bar_sync2<BAR_FULL, C::THREADS>(b);
            /* ---- FIR: one pass over the thread's NB+R-1 blocks, newest first; block m (relative to
             * block R*tid) feeds symbol ka+r with tap group bb = r - m + NB - 1 */
            const ulonglong2* rbase = reinterpret_cast<const ulonglong2*>(s_v + b * C::PCHUNK) + (C::PADW + 1) * tid;
            f32x2 acc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = 0ull; /* (+0.0f, +0.0f) */
            if (have_any) {
                f32x2 cv[8], nv[8];
                auto load_block = [&](f32x2* dst, int m) {
                    /* chunks 4R*tid + 4m + j  ->  position (4R+1)*tid + x + x/PADW, x = 4m + j */
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int x = 4 * m + j;
                        const ulonglong2 t = rbase[x + x / C::PADW];
                        dst[2 * j + 0] = t.x;
                        dst[2 * j + 1] = t.y;
                    }
                };
                load_block(cv, C::NSTEP - 1);
#pragma unroll
                for (int s = 0; s < C::NSTEP; ++s) {
                    const int m = C::NSTEP - 1 - s;
                    if (s + 1 < C::NSTEP) load_block(nv, m - 1); /* one block ahead of its use */
#pragma unroll
                    for (int e = 7; e >= 0; --e) {
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const int bb = r - m + C::NB - 1;
                            const int t = 8 * bb + 7 - OFF - e; /* tap index, ascending as e descends */
                            if (bb >= 0 && bb < C::NB && t >= 0 && t < NT)
                                acc[r] = mac2<FMA>(acc[r], cv[e], pk2(taps.hh[t].x, taps.hh[t].y), one);
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) cv[e] = nv[e];
                }
            }
            /* every shared load of buffer b has returned (its data fed the MACs above) */
            if (i + 2 < nframes) named_bar_arrive2<BAR_EMPTY, C::THREADS>(b);

            /* ---- slice, pack, count */
            if (vec_out) {
                uint32_t symw = 0, bitw[2] = {0u, 0u}, nerr = 0;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float2 y = unpk2(acc[r]);
                    const float I = __fmul_rn(a.rx_gain, y.x), Q = __fmul_rn(a.rx_gain, y.y);
                    const uint32_t s = slice_point4(s_slut + toff[r], a.n_const, I, Q);
                    symw |= s << (8 * r);
                    bitw[r / 2] |= ((s >> 1) | ((s & 1u) << 8)) << (16 * (r % 2));
                    if (a.soft) a.soft[orow + r] = make_float2(I, Q);
                    if (a.ref_bits) {
                        uint32_t ref;
                        if (ref_vec) {
                            const uint32_t wd = (r / 2) ? refw.y : refw.x;
                            const uint32_t h = wd >> (16 * (r % 2));
                            ref = ((h & 1u) << 1) | ((h >> 8) & 1u);
                        } else {
                            ref = pack_symbol(refp + 2 * r, 2);
                        }
                        nerr += __popc(ref ^ s);
                    }
                }
                if (a.sym) {
                    if (R == 4) *reinterpret_cast<uint32_t*>(a.sym + orow) = symw;
                    else *reinterpret_cast<uint16_t*>(a.sym + orow) = (uint16_t)symw;
                }
                if (a.bits) {
                    if (R == 4) *reinterpret_cast<uint2*>(a.bits + 2 * orow) = make_uint2(bitw[0], bitw[1]);
                    else *reinterpret_cast<uint32_t*>(a.bits + 2 * orow) = bitw[0];
                }
                if (a.ref_bits) {
                    err += nerr;
                    cmp += 2 * R;
                }
            } else if (have_any) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (ka + r < a.K) {
                        const float2 y = unpk2(acc[r]);
                        const float I = __fmul_rn(a.rx_gain, y.x), Q = __fmul_rn(a.rx_gain, y.y);
                        const uint32_t s = slice_point4(s_slut + toff[r], a.n_const, I, Q);
                        err += emit_symbol(a, f0 + i, ka + r, s, I, Q);
                        cmp += a.ref_bits ? a.bps : 0u;
                    }
                }
            }
            if (refp) refp += a.ref_stride;
        }
    }
    block_count(a, err, cmp);
}

/* ------------------------------------------------------------------ host side */
template <int NT, int OFF, bool FMA, bool NOISE, int R, int CW, int PW, int MINB>
cudaError_t rx_ws_launch_t(const RxArgs& a, const float* h_taps, cudaStream_t stream)
{
    using C = RxWsCfg<NT, OFF, R, CW, PW>;
    dim3 grid((unsigned)((a.K + C::TS - 1) / C::TS), (unsigned)((a.F + a.frames_per_block - 1) / a.frames_per_block));
    const TapsParam<NT> tp = make_taps_param<NT>(h_taps);
    const size_t smem = C::smem(a.n_tables * a.n_const);
    auto kern = rx_ws_kernel<NT, OFF, FMA, NOISE, R, CW, PW, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    kern<<<grid, C::THREADS, smem, stream>>>(a, tp);
    return cudaGetLastError();
}

/* all (OFF, FMA, NOISE) combinations of one shape */
template <int NT, int R, int CW, int PW, int MINB>
cudaError_t rx_ws_dispatch(const RxArgs& a, const float* h_taps, bool fma, cudaStream_t stream)
{
    const bool odd = (a.delay & 1u) != 0; /* OFF = 0 for odd delay, 1 for even */
    const bool noise = a.nz.sigma != 0.0f;
#define MG_WS_CASE(O, F, N) return rx_ws_launch_t<NT, O, F, N, R, CW, PW, MINB>(a, h_taps, stream)
    if (odd) {
        if (fma) { if (noise) MG_WS_CASE(0, true, true); else MG_WS_CASE(0, true, false); }
        else     { if (noise) MG_WS_CASE(0, false, true); else MG_WS_CASE(0, false, false); }
    } else {
        if (fma) { if (noise) MG_WS_CASE(1, true, true); else MG_WS_CASE(1, true, false); }
        else     { if (noise) MG_WS_CASE(1, false, true); else MG_WS_CASE(1, false, false); }
    }
#undef MG_WS_CASE
}

} /* namespace mg */
