/*
 * rx_fast.cuh -- fast decimating RX for 8 samples per symbol and a compile-time tap count NT:
 * the hot kernel of the loopback (replaces demodulator.rs:44-55 + two fir.rs:18-34 filters
 * of the reference, plus the decimator / slicer / error-count extension).
 *
 * The tile is a run of 8-sample BLOCKS; block B holds tile-local samples [8B, 8B+7].  The
 * decision instant of tile symbol r sits at position 7-OFF of block r+NB-1 (OFF = 0 for an
 * odd decision delay, 1 for an even one, so that blocks coincide with the 16-byte aligned
 * sample pairs global memory is read in), and tap i of symbol r reads element e of block
 * r+NB-1-b with i = 8b + 7 - OFF - e.
 *
 *  phase A  the tile's global loads (128-bit, two complex samples; the next frame's tile was
 *           prefetched towards L2 during the previous FIR) are all issued before the first
 *           use; the real parts are multiplied by the CTA-resident (cos, -sin) table and
 *           stored as interleaved (vi, vq) pairs: one 128-bit shared store per pair;
 *  phase B  every thread owns R consecutive symbols and walks its NB+R-1 blocks ONCE, newest
 *           first: a block is loaded into registers one step ahead (conflict-free 128-bit
 *           shared loads at thread_base + compile-time offset) and feeds all R symbols, each
 *           with its own compile-time tap group, so every accumulator still sees taps
 *           0..NT-1 in order; taps are constant-bank operands.  Shared traffic of the FIR is
 *           (NB+R-1)*64 B per R symbols: 22 B/sample at R = 4, NT = 64;
 *  phase C  slice, pack, count errors; one R-byte symbol store and one 2R-byte bit store per thread.
 *
 *  TXF      (template flag) the fused loopback: phase A makes the tile's TX samples from the frame's bits instead of
 *           loading them, stores them once, and mixes them from registers -- see the block comment in the kernel.
 *
 * Shared layout: 16-byte chunk c (samples 2c, 2c+1) lives at chunk position c + c/(4R): one
 * chunk of padding per thread stride, so the 8 lanes of a quarter warp (stride 4R chunks in
 * phase B, stride 1 in phase A) always hit 8 distinct bank groups.
 */
#pragma once

#include <atomic>
#include <type_traits>

#include "common.cuh"

namespace mg {

#ifndef MG_TXF_PACKED_MIX
#define MG_TXF_PACKED_MIX 1
#endif
#ifndef MG_RX_LDMODE
#define MG_RX_LDMODE 3
#endif
/* L1 policy of the two per-frame load streams (measured on B200, C2): 0 = 0.583 ms,
 * x with L1::no_allocate = 0.70-0.71 ms (much worse), see DESIGN.md */
#if MG_RX_LDMODE == 0
#define LDX(p) __ldcs(p)
#define LDC(p) __ldg(p)
#elif MG_RX_LDMODE == 3
#define LDX(p) __ldg(p)
#define LDC(p) __ldg(p)
#elif MG_RX_LDMODE == 4
#define LDX(p) __ldcs(p)
#define LDC(p) ld_keep_f4(p)
#elif MG_RX_LDMODE == 5
#define LDX(p) __ldg(p)
#define LDC(p) ld_keep_f4(p)
#else
#define LDX(p) ld_stream_f4(p)
#define LDC(p) __ldg(p)
#endif

/* ---- TMEM as a per-CTA scratchpad for frame-invariant NCO values (TMC > 0) -------------------------
 * Every thread stages the SAME tile chunks for every frame, so their (cos, sin) never change.  The first
 * TMC/4 chunks of each thread are parked in tensor memory (one 32-bit column per value, the thread's own
 * lane) and read back with tcgen05.ld each frame: that traffic bypasses the L1 / shared data pipe, which is
 * what bounds this kernel.  32 columns per CTA x 12 CTAs per SM fit the 512-column budget. */
template <int NCOL>
__device__ __forceinline__ uint32_t tmem_alloc(uint32_t* smem_slot)
{
    if ((threadIdx.x >> 5) == 0) { /* one whole warp allocates, then lets other CTAs allocate */
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_slot)), "n"(NCOL) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *smem_slot;
}
template <int NCOL>
__device__ __forceinline__ void tmem_free(uint32_t taddr)
{
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOL) : "memory");
}
/* this warp's lane quarter of the allocation: lane field (bits 31..16) = 32 * (warp % 4) */
__device__ __forceinline__ uint32_t tmem_warp_addr(uint32_t taddr) { return taddr + ((uint32_t)((threadIdx.x >> 5) & 3) << 21); }
__device__ __forceinline__ void tmem_st32(uint32_t addr, const float* v)
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(addr),
        "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]),
        "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]), "f"(v[20]),
        "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]),
        "f"(v[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, float* v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,"
        "%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
          "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]),
          "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]),
          "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
        : "r"(addr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int NT, int OFF, int THREADS, int R>
struct RxFastCfg {
    static constexpr int NB = (NT + OFF + 7) / 8; /* blocks that reach one symbol */
    static constexpr int TS = R * THREADS;        /* symbols per tile */
    static constexpr int NBLK = TS + NB - 1;      /* blocks staged per tile */
    static constexpr int NSAMP = NBLK * 8;
    static constexpr int NCHUNK = NSAMP / 2;      /* 16-byte chunks */
    static constexpr int PADW = 4 * R;            /* one padding chunk per PADW chunks */
    static constexpr int PCHUNK = NCHUNK + NCHUNK / PADW + 1;
    static constexpr int ITER = (NCHUNK + THREADS - 1) / THREADS;
    static constexpr int NSTEP = NB + R - 1;      /* blocks one thread walks */
    static constexpr size_t SMEM_V = sizeof(float4) * PCHUNK;
    static size_t smem(uint32_t lut_entries) { return SMEM_V + sizeof(float2) * lut_entries; }
    static_assert(THREADS % PADW == 0, "phase A position arithmetic needs THREADS % (4R) == 0");
};

/* nearest point over a table in shared memory, four candidates per trip */
__device__ __forceinline__ uint32_t slice_point4(const float2* t, uint32_t n, float I, float Q)
{
    if (n < 4) return slice_point(t, n, I, Q);
    uint32_t best = 0;
    float bd = 0.0f;
    for (uint32_t s = 0; s < n; s += 4) {
        float d[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 c = t[s + j];
            const float di = __fsub_rn(I, c.x), dq = __fsub_rn(Q, c.y);
            d[j] = __fadd_rn(__fmul_rn(di, di), __fmul_rn(dq, dq));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if ((s + j == 0) || d[j] < bd) {
                bd = d[j];
                best = s + j;
            }
    }
    return best;
}

/* the same search for a 4-point table held in registers (QPSK-sized constellations): identical float
 * operations and tie rule (strict <, lowest index wins), no shared loads, no loop */
__device__ __forceinline__ uint32_t slice_point_reg4(const float2 (&t)[4], float I, float Q)
{
    float d[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float di = __fsub_rn(I, t[j].x), dq = __fsub_rn(Q, t[j].y);
        d[j] = __fadd_rn(__fmul_rn(di, di), __fmul_rn(dq, dq));
    }
    uint32_t best = 0;
    float bd = d[0];
#pragma unroll
    for (int j = 1; j < 4; ++j)
        if (d[j] < bd) {
            bd = d[j];
            best = j;
        }
    return best;
}

/* (cos, sin) of an NCO angle plus offset as ONE out-of-line routine: the raw-wire variant evaluates it 34 times per thread and
 * frame; inlined, those copies made the kernel 5000 instructions long and it stalled on instruction fetch (no_instruction was
 * its top stall: profiles) */
static __device__ __noinline__ float2 sincos_nco_call(float y)
{
    float s, c;
    mg_sincosf_nco(y, &s, &c);
    return make_float2(c, s);
}

template <int NT, int OFF, bool FMA, bool NOISE, int THREADS, int MINB, int R, int PF, int TMC, bool TXF = false>
__global__ void __launch_bounds__(THREADS, MINB)
    rx_fast_kernel(const __grid_constant__ RxArgs a, const __grid_constant__ TapsParam<NT> taps)
{
    using C = RxFastCfg<NT, OFF, THREADS, R>;
    static_assert((THREADS & (THREADS - 1)) == 0, "tid is masked to THREADS - 1 so that ptxas can drop the per-chunk range checks");
    static_assert(C::ITER <= 64, "one validity bit per staged chunk");
    using vmask_t = typename std::conditional<(C::ITER > 32), unsigned long long, uint32_t>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ulonglong2* s_v = reinterpret_cast<ulonglong2*>(smem_raw); /* padded chunks of packed pairs (vi0,vq0), (vi1,vq1) */
    float2* s_slut = reinterpret_cast<float2*>(s_v + C::PCHUNK);

    const int tid = threadIdx.x & (THREADS - 1);
    for (uint32_t i = tid; i < a.n_tables * a.n_const; i += THREADS) s_slut[i] = a.slut[i];

    /* tile-major launch order (blockIdx.x = frame group) makes the CTAs that are resident together work on
     * the same sample tile, so they share its slice of the NCO table in L1 */
    const u64 f0 = (u64)(a.tile_major ? blockIdx.x : blockIdx.y) * a.frames_per_block;
    const u64 f1 = min(a.F, f0 + a.frames_per_block);
    const u64 k0 = (u64)(a.tile_major ? blockIdx.y : blockIdx.x) * C::TS;
    /* sample index of tile-local j = 0; even by the choice of OFF */
    const long long nbase = (long long)(k0 * 8 + a.delay) + OFF - 8 * C::NB + 1;
    /* the sample range of the tile that exists in the frame, [vlo, vhi) as tile-local pair indices: L and
     * nbase are even, so a 16-byte pair is either entirely inside the frame or entirely outside */
    const long long vlo_n = nbase < 0 ? 0 : nbase;
    const long long vhi_n = (u64)(nbase + C::NSAMP) > a.L ? (long long)a.L : nbase + C::NSAMP;
    /* INTERIOR tiles (every staged sample exists in the frame: all but the first and the last tile of a frame) run a
     * phase A without any per-chunk validity test; the noisy and the register-prefetch forms always take the edge form */
    /* RAW (PF 7 = f32 rows, PF 8 = i16 rows): the real-valued wire of src/bin/demodulate.rs:29-43 with a phase offset that may
     * differ from frame to frame (Demodulator::lock_phase leaves one in every frame's PLL).  The frame-invariant part of the
     * NCO angle, phase(n) of carrier.rs:17-19, is what gets parked in tensor memory; every frame adds its own offset and
     * evaluates cos / sin (libm_f32.h: glibc's, bit for bit) while loading -- demodulator.rs:51-54 as written.  No table, no
     * prefetch; always the edge form of phase A. */
    constexpr int RAW = PF == 7 ? 1 : (PF == 8 ? 2 : 0);
    static_assert(RAW == 0 || (!TXF && !NOISE && TMC > 0), "the raw-wire variant parks its phases in tensor memory");
    const bool edge = NOISE || PF == 2 || RAW != 0 || nbase < 0 || (u64)(nbase + C::NSAMP) > a.L;

    /* NCO values of the tile: read every frame from the zero-padded global table (ChannelView::cs_tab).
     * The tile's 17 KB slice stays L1/L2-resident across the frame loop, costs no shared memory and no
     * setup pass; cs4[q] = (cos, sin) of samples nbase+2q, nbase+2q+1 (zero outside the frame). */
    const float4* cs4 = RAW ? nullptr : reinterpret_cast<const float4*>(chan_table(a.ch, f0) + nbase) + tid;
    const float raw_w = RAW ? chan_w(a.ch, f0) : 0.0f;
    /* phase(n) of tile-local chunk q's two samples; outside the frame the value is never used (validity mask) */
    auto raw_phase = [&](int q, float* p0, float* p1) {
        const long long n = nbase + 2 * (long long)q;
        *p0 = n >= 0 ? nco_phase(raw_w, a.sample0 + (u64)n) : 0.0f;
        *p1 = n + 1 >= 0 ? nco_phase(raw_w, a.sample0 + (u64)(n + 1)) : 0.0f;
    };
    /* TMC columns of tensor memory per thread hold the NCO values of its first TMC/4 chunks */
    constexpr int TCH = TMC / 4;
    __shared__ uint32_t s_tmem;
    uint32_t taddr = 0, twarp = 0;
    if (TMC > 0) {
        static_assert(TMC == 0 || TMC == 32 || TMC == 64 || TMC == 128, "one, two or four tcgen05 32x32b.x32 transfers per thread");
        static_assert(TMC == 0 || THREADS <= 128, "a CTA reaches TMEM lanes 32*(warp % 4)");
        taddr = tmem_alloc<(TMC > 0 ? TMC : 32)>(&s_tmem);
        twarp = tmem_warp_addr(taddr);
#pragma unroll
        for (int h = 0; h < TMC / 32; ++h) {
            float park[32];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                float4 t = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (RAW) raw_phase((8 * h + it) * THREADS + tid, &t.x, &t.y); /* (phase 0, phase 1, -, -) */
                else t = __ldg(cs4 + (8 * h + it) * THREADS); /* padded table: always readable */
                park[4 * it] = t.x; park[4 * it + 1] = t.y; park[4 * it + 2] = t.z; park[4 * it + 3] = t.w;
            }
            tmem_st32(twarp + 32 * h, park); /* column field = bits 15..0 */
        }
    }

    const u64 ka = k0 + (u64)R * tid; /* this thread's symbols: ka .. ka+R-1 */
    uint32_t toff[R];
#pragma unroll
    for (int r = 0; r < R; ++r) toff[r] = (uint32_t)((ka + r) % a.n_tables) * a.n_const;
    const bool have_any = ka < a.K, have_all = ka + R <= a.K;
    /* vector emit: every frame's symbol row must start on a multiple of R symbols */
    /* vector emit: every frame's symbol row must start on a multiple of VG symbols (R = 8 emits two groups of 4, so that
     * K = 8188 symbols per frame -- a multiple of 4, not of 8 -- still takes the vector path) */
    constexpr int VG = R == 8 ? 4 : R;
    const bool vec_out = a.bps == 2 && have_all && (a.K % VG == 0) &&
                         ((reinterpret_cast<uintptr_t>(a.sym) % VG) == 0) &&
                         ((reinterpret_cast<uintptr_t>(a.bits) % (2 * VG)) == 0);
    const bool ref_vec = a.ref_bits && vec_out && (a.ref_stride % (2 * VG) == 0) &&
                         ((reinterpret_cast<uintptr_t>(a.ref_bits) % (2 * VG)) == 0);

    /* one 4-point table for every symbol: keep it in registers */
    const bool lut4 = a.n_const == 4 && a.n_tables == 1;
    float2 rl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) rl[j] = lut4 ? a.slut[j] : make_float2(0.0f, 0.0f);

    uint32_t err = 0, cmp = 0;
    const int wbase = tid + tid / C::PADW;              /* phase A: chunk position of chunk `tid` */
    const ulonglong2* rbase = s_v + (C::PADW + 1) * tid; /* phase B: chunk position of chunk 4R*tid */
    const f32x2 one = pk2(taps.one.x, taps.one.y);
    const float2* frame = a.rx + f0 * a.L;
    u64 orow = f0 * a.K + ka; /* output index of symbol ka in the current frame */
    const uint8_t* refp = a.ref_bits ? a.ref_bits + f0 * a.ref_stride + ka * 2 : nullptr;
    float xr[C::ITER][2]; /* real parts of the tile being staged (PF == 2: of the NEXT frame, in flight during the FIR) */
    /* which of this thread's chunks exist in the frame (edge tiles): one bit per step, fixed for the whole frame loop, so
     * the loads below are predicated, never branched around (a branch would make ptxas drain outstanding
     * loads at the join and defeat any prefetch) */
    vmask_t vmask = 0;
    vmask_t vmask1 = 0; /* RAW: does the chunk's SECOND sample exist?  (a raw frame may hold an odd number of samples) */
#pragma unroll
    for (int it = 0; it < C::ITER; ++it) {
        const long long n = nbase + 2 * (it * THREADS + tid);
        if (it * THREADS + tid < C::NCHUNK && n >= vlo_n && n < vhi_n) vmask |= (vmask_t)1 << it;
        if (RAW && it * THREADS + tid < C::NCHUNK && n >= vlo_n && n + 1 < vhi_n) vmask1 |= (vmask_t)1 << it;
    }
    /* does step `it` stage a chunk at all?  Only the last step is partial, and by a compile-time thread count */
    auto staged = [&](int it) { return (it + 1) * THREADS <= C::NCHUNK || tid < C::NCHUNK - it * THREADS; };
    u64 raw_row = RAW ? f0 * a.raw_stride + a.raw_skip : 0; /* element index of the current frame's sample 0 in a.raw */
    auto load_tile = [&](const float2* fr, auto edge_c) {
        constexpr bool EDGE = decltype(edge_c)::value;
        if (RAW) { /* one load per sample: rows of the wire may start on any element (odd strides, odd preambles) */
            const long long e0 = (long long)raw_row + nbase + 2 * tid;
#pragma unroll
            for (int it = 0; it < C::ITER; ++it) {
                const long long e = e0 + 2 * it * THREADS;
                float x0 = 0.0f, x1 = 0.0f;
                if (RAW == 1) {
                    const float* src = reinterpret_cast<const float*>(a.raw);
                    if (((vmask >> it) & 1) != 0) x0 = __ldg(src + e);
                    if (((vmask1 >> it) & 1) != 0) x1 = __ldg(src + e + 1);
                } else { /* `x as f32` of an i16 sample (demodulate.rs:29) */
                    const short* src = reinterpret_cast<const short*>(a.raw);
                    /* the integer rides in the float register until the sample is used: converting here would make every thread wait
                     * for its loads at the top of phase A (long_scoreboard was 2.6 stall cycles per instruction that way) */
                    if (((vmask >> it) & 1) != 0) x0 = __int_as_float((int)__ldg(src + e));
                    if (((vmask1 >> it) & 1) != 0) x1 = __int_as_float((int)__ldg(src + e + 1));
                }
                xr[it][0] = x0;
                xr[it][1] = x1;
            }
            return;
        }
        const float4* src = reinterpret_cast<const float4*>(fr + nbase) + tid;
#pragma unroll
        for (int it = 0; it < C::ITER; ++it) {
            float4 t = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (EDGE ? ((vmask >> it) & 1) != 0 : staged(it)) t = LDX(src + it * THREADS);
            xr[it][0] = t.x;
            xr[it][1] = t.z;
        }
    };
    /* ---- TXF: the fused loopback.  The tile's TX samples are not read from memory but MADE here, from the frame's
     * bits (a.ref_bits, two bytes per symbol) with the arithmetic of the rectangular-hold TX kernel
     * (data.rs:66-79 hold, digital/qpsk.rs:23-35 as a 4-entry table, modulator.rs:37-48 mix: re = i*cos - q*sin,
     * im = i*sin + q*cos, each operation rounded on its own), stored to a.tx_out exactly once (a tile owns the samples
     * behind its halo; the first tile also owns the frame's head) and fed to the demodulator from registers: the
     * 8 B/sample read of the RX side disappears, the loopback is one pass of 8 B/sample written + bits.  Same NCO values
     * on both sides (no phase offset: the launcher checks), so every buffer is bit-identical to the two-kernel path. */
    const bool first_tile = k0 == 0;
    const bool store_tx = TXF && a.tx_out != nullptr;
    static_assert(!TXF || THREADS % 4 == 0, "symbol stride per step = 2*THREADS/8");
    /* The tile's symbols reach the threads through a small shared table of (i, q) PAIRS, one per symbol the tile touches:
     * every thread fetches one 8-byte word (4 symbols of 2 bit bytes) of the NEXT frame's row when the FIR starts (two
     * registers ride through the FIR), maps its 4 symbols through the constellation table (a.tx_iq, a constant-bank
     * lookup by the symbol index 2*b0 + b1: first byte = MSB, digital/util.rs:5-11) when the FIR is done and parks the
     * pairs in shared memory; phase A of the next frame then needs ONE 8-byte shared load per staged chunk and no bit
     * arithmetic.  (Round 1: 2-byte loads and select/sign logic per CHUNK, 8 instructions x 17 chunks per thread and
     * frame.)  Rows start on 8-byte boundaries (the launcher checks). */
    constexpr int BWORDS = (2 * C::NBLK + 7 + 7) / 8;
    static_assert(!TXF || BWORDS <= 3 * THREADS, "three words per thread at most");
    constexpr bool W3 = TXF && BWORDS > 2 * THREADS; /* a few threads carry a third word (R = 8 shapes) */
    __shared__ __align__(16) f32x2 s_sq[TXF ? 4 * BWORDS : 2];
    const long long brow0 = (2 * (nbase >> 3)) & ~7ll; /* row byte offset of table entry 0; negative in the first tile */
    /* this thread's first chunk lies in symbol (nbase + 2*tid) >> 3; chunk it*THREADS + tid: THREADS/4 symbols on per step */
    const int sqoff = TXF ? (int)(((nbase + 2 * (long long)tid) >> 3) - (brow0 >> 1)) : 0;
    const uint8_t* brow = TXF ? a.ref_bits + f0 * a.ref_stride : nullptr;
    float4* txrow = TXF ? reinterpret_cast<float4*>(a.tx_out + f0 * a.L + nbase) + tid : nullptr;
    unsigned long long nb0 = 0, nb1 = 0, nb2 = 0;
    const long long bo0 = brow0 + 8 * tid, bo1 = bo0 + 8 * THREADS, bo2 = bo1 + 8 * THREADS;
    const bool bw2 = W3 && tid + 2 * THREADS < BWORDS;
    const bool bv2 = bw2 && bo2 >= 0 && bo2 + 8 <= (long long)a.ref_stride;
    const bool bw0 = TXF && tid < BWORDS, bw1 = TXF && tid + THREADS < BWORDS; /* has a first / second word slot */
    const bool bv0 = bw0 && bo0 >= 0 && bo0 + 8 <= (long long)a.ref_stride;     /* ... and it lies inside the row */
    const bool bv1 = bw1 && bo1 >= 0 && bo1 + 8 <= (long long)a.ref_stride;
    auto fetch_bits = [&](const uint8_t* row) {
        nb0 = nb1 = 0; /* symbols outside the frame map to entry 0: finite, and multiplied by the NCO table's zeros */
        if (bv0) nb0 = __ldg(reinterpret_cast<const unsigned long long*>(row + bo0));
        if (bv1) nb1 = __ldg(reinterpret_cast<const unsigned long long*>(row + bo1));
        if (W3) {
            nb2 = 0;
            if (bv2) nb2 = __ldg(reinterpret_cast<const unsigned long long*>(row + bo2));
        }
    };
    auto map_word = [&](unsigned long long w, f32x2* dst) {
        f32x2 e[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t h = (uint32_t)(w >> (16 * j));
            const float2 p = a.tx_iq[((h & 1u) << 1) | ((h >> 8) & 1u)];
            e[j] = pk2(p.x, p.y);
        }
        reinterpret_cast<ulonglong2*>(dst)[0] = make_ulonglong2(e[0], e[1]);
        reinterpret_cast<ulonglong2*>(dst)[1] = make_ulonglong2(e[2], e[3]);
    };
    auto park_bits = [&]() {
        if (bw0) map_word(nb0, s_sq + (TXF ? 4 * tid : 0));
        if (bw1) map_word(nb1, s_sq + (TXF ? 4 * (tid + THREADS) : 0));
        if (W3 && bw2) map_word(nb2, s_sq + (W3 ? 4 * (tid + 2 * THREADS) : 0));
    };
    if (TXF && f0 < f1) {
        fetch_bits(brow);
        park_bits();
    }
    /* ---- phase A: load the tile (all loads issued before the first use), mix, stage.  EDGE = the tile reaches past a
     * frame end (validity bits per chunk), else every chunk exists and only the last step is partial */
    auto phase_a = [&](const float2* fr, u64 f, auto edge_c) {
        constexpr bool EDGE = decltype(edge_c)::value;
        const float raw_po = RAW ? chan_po(a.ch, f) : 0.0f; /* this frame's PLL.phase_offset (or the call's / the channel's) */
        if (!TXF && PF != 2) load_tile(fr, edge_c);
        if (NOISE) {
            /* AWGN on the fly (oracle/modem_oracle.h "AWGN"): one Philox block serves an aligned QUAD of samples = the chunks
             * of two neighbouring lanes, so the lanes of a pair split the generator work -- over two steps the even lane runs
             * the block of the first step, the odd lane the block of the second, and each hands the other the two words it
             * does not need itself (two shuffles): one Philox per TWO chunks instead of one per chunk.  Four steps per trip
             * give every lane two independent Philox chains and four independent Box-Muller evaluations to interleave; the
             * trip count stays rolled to keep the code in the instruction cache.  When the tile's quads do not start on an
             * even lane (nbase = 2 mod 4: some even decision delays) every lane runs its own block. */
            constexpr int G = 4;
            const u64 gf = a.nz.frame0 + f;
            const bool odd = (tid & 1) != 0;
            const bool paired = ((nbase >> 1) & 1) == 0;
#pragma unroll 1
            for (int it0 = 0; it0 < C::ITER; it0 += G) {
                uint32_t w0[G], w1[G]; /* the two words of this lane's chunk at step it0 + g */
                if (paired) {
#pragma unroll
                    for (int h = 0; h < G / 2; ++h) {
                        const int itm = it0 + 2 * h + (odd ? 1 : 0); /* the step whose block this lane runs */
                        const long long n = nbase + 2 * ((long long)itm * THREADS + (tid & ~1));
                        uint32_t r[4];
                        noise_quad(a.nz, gf, (u64)(n >> 2), 0, r);
                        const uint32_t g0 = __shfl_xor_sync(0xffffffffu, odd ? r[0] : r[2], 1);
                        const uint32_t g1 = __shfl_xor_sync(0xffffffffu, odd ? r[1] : r[3], 1);
                        w0[2 * h] = odd ? g0 : r[0];     /* step it0 + 2h: the even lane's block; the odd lane's chunk = words 2, 3 */
                        w1[2 * h] = odd ? g1 : r[1];
                        w0[2 * h + 1] = odd ? r[2] : g0; /* step it0 + 2h + 1: the odd lane's block */
                        w1[2 * h + 1] = odd ? r[3] : g1;
                    }
                } else {
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const long long n = nbase + 2 * ((long long)(it0 + g) * THREADS + tid);
                        uint32_t r[4];
                        noise_quad(a.nz, gf, (u64)(n >> 2), 0, r);
                        w0[g] = (n & 2) ? r[2] : r[0];
                        w1[g] = (n & 2) ? r[3] : r[1];
                    }
                }
                float n0[G], n1[G];
#pragma unroll
                for (int g = 0; g < G; ++g) box_muller(w0[g], w1[g], &n0[g], &n1[g]);
                /* xr[] must be indexed statically to stay in registers: one switch case per trip (a predicated select over
                 * all ITER x G combinations cost 240 instructions per trip for 16 useful ones) */
                auto apply = [&](auto trip_c) {
                    constexpr int T = decltype(trip_c)::value;
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        constexpr int JMAX = C::ITER - 1;
                        const int j = T * G + g < JMAX ? T * G + g : JMAX; /* clamp keeps the index in range for dead iterations */
                        if (T * G + g < C::ITER && ((vmask >> j) & 1) != 0) {
                            xr[j][0] = __fadd_rn(xr[j][0], __fmul_rn(a.nz.sigma, n0[g]));
                            xr[j][1] = __fadd_rn(xr[j][1], __fmul_rn(a.nz.sigma, n1[g]));
                        }
                    }
                };
                static_assert(!NOISE || C::ITER <= 8 * G, "one switch case per trip");
                switch (it0 / G) {
                case 0: apply(std::integral_constant<int, 0>{}); break;
                case 1: apply(std::integral_constant<int, 1>{}); break;
                case 2: apply(std::integral_constant<int, 2>{}); break;
                case 3: apply(std::integral_constant<int, 3>{}); break;
                case 4: apply(std::integral_constant<int, 4>{}); break;
                case 5: apply(std::integral_constant<int, 5>{}); break;
                case 6: apply(std::integral_constant<int, 6>{}); break;
                default: apply(std::integral_constant<int, 7>{}); break;
                }
            }
        }
        /* the parked NCO values come back 32 columns (8 steps) at a time, each batch when the previous one is used up */
        float parked[TMC > 0 ? 32 : 4];
#pragma unroll
        for (int it = 0; it < C::ITER; ++it) {
            if (TMC > 0 && it < TCH && it % 8 == 0) tmem_ld32(twarp + 4 * it, parked);
            if (staged(it)) {
                float4 cs;
                if (RAW) { /* demodulator.rs:51: phase = carrier.next() + pll.phase_offset; :53-54 its cos and sin */
                    float p0, p1;
                    if (it < TCH) { p0 = parked[(4 * it) % 32]; p1 = parked[(4 * it + 1) % 32]; }
                    else raw_phase(it * THREADS + tid, &p0, &p1);
                    cs = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    /* measured: the f32 wire is faster with the out-of-line routine (1.99 -> 1.68 ms for the bench's 4096 frames: no
                     * more instruction-fetch stalls), the i16 wire -- whose 2-byte loads arrive late and want the longer inlined
                     * stretch of work in front of their first use -- with the inlined one (2.02 against 2.21 ms) */
                    if (RAW == 1) {
                        if (((vmask >> it) & 1) != 0) { const float2 t = sincos_nco_call(__fadd_rn(p0, raw_po)); cs.x = t.x; cs.y = t.y; }
                        if (((vmask1 >> it) & 1) != 0) { const float2 t = sincos_nco_call(__fadd_rn(p1, raw_po)); cs.z = t.x; cs.w = t.y; }
                    } else {
                        if (((vmask >> it) & 1) != 0) mg_sincosf_nco(__fadd_rn(p0, raw_po), &cs.y, &cs.x);
                        if (((vmask1 >> it) & 1) != 0) mg_sincosf_nco(__fadd_rn(p1, raw_po), &cs.w, &cs.z);
                    }
                } else {
                    cs = (TMC > 0 && it < TCH) ? make_float4(parked[(4 * it) % 32], parked[(4 * it + 1) % 32], parked[(4 * it + 2) % 32], parked[(4 * it + 3) % 32])
                                               : LDC(cs4 + it * THREADS);
                }
                float x0r, x1r;
                if (TXF) {
                    /* (i, q) of the chunk's symbol.  (Fetching these pairs a few steps ahead, or all up front, measured
                     * 4-10 % SLOWER at this CTA shape although 19 % of the stall samples sit on this load: the extra live
                     * registers cost more than the round trip, which the other three warps of the scheduler cover.) */
                    const float2 sq = unpk2(s_sq[TXF ? sqoff + (THREADS / 4) * it : 0]);
                    /* modulator.rs:37-43 on packed pairs: (i*c, i*s) and (q*s, q*c) by two FMUL2, then
                     * (i*c - q*s, i*s + q*c) by one FFMA2 with (-1, +1): a product by +-1 is exact, so the fma's one
                     * rounding is the rounding of the reference's subtraction / addition */
                    const f32x2 ii = pk2(sq.x, sq.x), qq = pk2(sq.y, sq.y), pm = pk2(-1.0f, 1.0f);
                    const f32x2 x0 = fma2(mul2(qq, pk2(cs.y, cs.x)), pm, mul2(ii, pk2(cs.x, cs.y)));
                    const f32x2 x1 = fma2(mul2(qq, pk2(cs.w, cs.z)), pm, mul2(ii, pk2(cs.z, cs.w)));
                    /* this tile owns local samples >= 8*(NB-1) (what the previous tile did not reach); outside the frame
                     * (edge tiles) nothing is stored */
                    const bool own = EDGE ? ((vmask >> it) & 1) != 0 && (first_tile || it > 0 || tid >= 4 * (C::NB - 1))
                                          : (it > 0 || tid >= 4 * (C::NB - 1));
                    if (own && store_tx) __stcs(reinterpret_cast<float4*>(txrow + it * THREADS), make_float4(unpk2(x0).x, unpk2(x0).y, unpk2(x1).x, unpk2(x1).y));
                    /* outside the frame the NCO table holds zeros: x is +-0 there, like the zero the unfused kernel stages */
                    x0r = unpk2(x0).x;
                    x1r = unpk2(x1).x;
                } else if (RAW == 2) {
                    /* `x as f32` of the i16 sample (demodulate.rs:29), exact and without a conversion instruction (those share the XU
                     * pipe with the conversions inside cos / sin): 1.5 * 2^23 + x is the binary32 with the bits 0x4B400000 + x */
                    x0r = __fsub_rn(__int_as_float(0x4B400000 + __float_as_int(xr[it][0])), 12582912.0f);
                    x1r = __fsub_rn(__int_as_float(0x4B400000 + __float_as_int(xr[it][1])), 12582912.0f);
                } else {
                    x0r = xr[it][0];
                    x1r = xr[it][1];
                }
                /* demodulator.rs:53-54 stages x*cos and x*(-sin).  The Q rail is kept NEGATED here, (x*cos, x*sin): one
                 * FMUL2 per sample with the NCO pair as it lies in the table (a lane-wise negation costs an extra
                 * instruction per sample), and phase C takes 0 - acc.  Exact: round-to-nearest is symmetric, so every
                 * product and every partial sum of the negated rail is the negation of the reference's, and a partial
                 * sum that is zero is +0 in both (the fold starts from +0 and x + (-x) = +0), which 0 - acc restores.
                 * chunk q = it*THREADS + tid sits at position q + q/PADW = wbase + it*(THREADS + THREADS/PADW) */
                s_v[wbase + it * (THREADS + THREADS / C::PADW)] =
                    make_ulonglong2(mul2(pk2(x0r, x0r), pk2(cs.x, cs.y)), mul2(pk2(x1r, x1r), pk2(cs.z, cs.w)));
            }
        }
    };
    if (PF == 2 && f0 < f1 && !TXF) load_tile(frame, std::true_type{});
    for (u64 f = f0; f < f1; ++f, orow += a.K, frame += a.L, raw_row += a.raw_stride) {
        __syncthreads(); /* previous frame's phase B finished; s_slut, s_sq visible */
        if (edge) phase_a(frame, f, std::true_type{});
        else phase_a(frame, f, std::false_type{});
        __syncthreads();
        if (TXF) {
            brow += a.ref_stride;
            txrow += a.L / 2;
            if (f + 1 < f1) fetch_bits(brow); /* the next frame's words fly during the FIR */
        }
        if (PF == 2 && f + 1 < f1 && !TXF) load_tile(frame + a.L, std::true_type{}); /* next frame's loads fly during the FIR */
        /* pull the next frame's tile towards L2 while the FIR runs: PF 1 = one prefetch per 128-byte
         * line through the LSU, PF 3 = one bulk (TMA) L2 prefetch of the whole tile by one thread */
        if (PF == 1 && f + 1 < f1 && !TXF) {
            const char* nxt = reinterpret_cast<const char*>(frame + a.L + vlo_n);
            for (int o = tid * 128; o < (int)(vhi_n - vlo_n) * 8; o += THREADS * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + o));
        }
        /* PF 5 = the next frame's tile is pulled INTO L1, one prefetch per 128-byte line, while the FIR runs: phase A of the next
         * frame then finds its samples ~40 cycles away instead of an L2 round trip (20 % of the 129-tap kernel's stall samples
         * sat on that wait).  Needs an L1 that holds the next tiles of all resident CTAs (8 x 17 KB): the launcher asks for a
         * shared-memory carve-out of RX_CARVEOUT % (164 KB for 8 CTAs, 92 KB of L1) instead of the maximum.  Measured at C2 / C3 /
         * C5: RX 0.488 -> 0.461, 2.66 -> 2.53, 0.500 -> 0.467 ms; with the maximum carve-out (28 KB of L1) it is slower than PF 3. */
        if (PF == 5 && f + 1 < f1 && !TXF) {
            const char* nxt = reinterpret_cast<const char*>(frame + a.L + vlo_n);
            for (int o = tid * 128; o < (int)(vhi_n - vlo_n) * 8; o += THREADS * 128)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(nxt + o));
        }
        if (PF == 3 && !TXF && f + 1 < f1 && tid == 0 && vhi_n > vlo_n) {
            const char* nxt = reinterpret_cast<const char*>(frame + a.L + vlo_n);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nxt), "r"((int)(vhi_n - vlo_n) * 8) : "memory");
        }
        uint32_t refw[4] = {0u, 0u, 0u, 0u}; /* 2R reference bits, one byte each */
        if (ref_vec) {
            if (R == 8) {
                const uint2 t = __ldg(reinterpret_cast<const uint2*>(refp)), u = __ldg(reinterpret_cast<const uint2*>(refp) + 1);
                refw[0] = t.x; refw[1] = t.y; refw[2] = u.x; refw[3] = u.y;
            } else if (R == 4) {
                const uint2 t = __ldg(reinterpret_cast<const uint2*>(refp));
                refw[0] = t.x; refw[1] = t.y;
            } else {
                refw[0] = __ldg(reinterpret_cast<const uint32_t*>(refp));
            }
        }

        /* ---- phase B: one pass over the thread's NB+R-1 blocks, newest first; block m (relative to
         * block R*tid) feeds symbol ka+r with tap group b = r - m + NB - 1.  (vi, vq) of a sample
         * is one packed pair; a MAC on both rails is FMUL2 + FFMA2 (common.cuh, f32x2). */
        f32x2 acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0ull; /* (+0.0f, +0.0f) */
        if (have_any) {
            f32x2 cv[8], nv[8];
            auto load_block = [&](f32x2* dst, int m) {
                /* chunks 4R*tid + 4m + i  ->  position (4R+1)*tid + x + x/PADW, x = 4m + i */
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int x = 4 * m + i;
                    const ulonglong2 t = rbase[x + x / C::PADW];
                    dst[2 * i + 0] = t.x;
                    dst[2 * i + 1] = t.y;
                }
            };
            load_block(cv, C::NSTEP - 1);
            /* products run one element AHEAD of their accumulation (pend[] holds fl(v*h) of the previous
             * element), so an FFMA2 never waits on the FMUL2 issued just before it; per accumulator the
             * order of additions is unchanged: taps 0..NT-1 */
            f32x2 pend[R];
            bool have_pend[R];
#pragma unroll
            for (int r = 0; r < R; ++r) have_pend[r] = false;
#pragma unroll
            for (int s = 0; s < C::NSTEP; ++s) {
                const int m = C::NSTEP - 1 - s;
                if (s + 1 < C::NSTEP) load_block(nv, m - 1); /* one block ahead of its use */
#pragma unroll
                for (int e = 7; e >= 0; --e) {
                    f32x2 prod[R];
                    bool have_prod[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int b = r - m + C::NB - 1;
                        const int i = 8 * b + 7 - OFF - e; /* tap index, ascending as e descends */
                        have_prod[r] = b >= 0 && b < C::NB && i >= 0 && i < NT;
                        if (have_prod[r]) {
                            if (FMA) acc[r] = fma2(cv[e], pk2(taps.hh[i].x, taps.hh[i].y), acc[r]);
                            else prod[r] = mul2(cv[e], pk2(taps.hh[i].x, taps.hh[i].y));
                        }
                    }
                    if (!FMA) {
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            if (have_pend[r]) acc[r] = fma2(acc[r], one, pend[r]);
                            have_pend[r] = have_prod[r];
                            if (have_prod[r]) pend[r] = prod[r];
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) cv[e] = nv[e];
            }
            if (!FMA) {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (have_pend[r]) acc[r] = fma2(acc[r], one, pend[r]);
            }
        }
        float ai[R], aq[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float2 t = unpk2(acc[r]);
            ai[r] = t.x;
            aq[r] = __fsub_rn(0.0f, t.y); /* the Q rail was accumulated negated (phase A) */
        }
        /* ---- phase C */
        if (vec_out) {
            uint32_t symw[2] = {0u, 0u}, bitw[4] = {0u, 0u, 0u, 0u};
            float I[R], Q[R];
            /* sign slicer (a.sign_slice: the launcher found the scaled table to be (-+A, -+B) in index order with
             * A, B within a factor 2 of each other): b0 = I > 0, b1 = Q > 0 IS the nearest-point search with its binary32
             * roundings and its tie rule whenever every |I|, |Q| of the thread lies in [ss_lo, ss_hi] = [min(A,B)/1024,
             * 4 min(A,B)] -- the true distance gap 4|I|A >= 2^-8 min^2 then exceeds every rounding error of the search
             * (< 2^-16 min^2) -- and anything else (tiny, huge, NaN) takes the search itself */
            bool quick = a.sign_slice != 0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                I[r] = __fmul_rn(a.rx_gain, ai[r]);
                Q[r] = __fmul_rn(a.rx_gain, aq[r]);
                quick = quick && fabsf(I[r]) >= a.ss_lo && fabsf(I[r]) <= a.ss_hi && fabsf(Q[r]) >= a.ss_lo && fabsf(Q[r]) <= a.ss_hi;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                uint32_t s;
                if (quick) s = ((~__float_as_uint(I[r]) >> 31) << 1) | (~__float_as_uint(Q[r]) >> 31);
                else s = lut4 ? slice_point_reg4(rl, I[r], Q[r]) : slice_point4(s_slut + toff[r], a.n_const, I[r], Q[r]);
                symw[r / 4] |= s << (8 * (r % 4));
                bitw[r / 2] |= ((s >> 1) | ((s & 1u) << 8)) << (16 * (r % 2));
                if (a.soft) a.soft[orow + r] = make_float2(I[r], Q[r]);
            }
            if (a.sym) {
                if (R == 8) {
                    reinterpret_cast<uint32_t*>(a.sym + orow)[0] = symw[0];
                    reinterpret_cast<uint32_t*>(a.sym + orow)[1] = symw[1];
                }
                else if (R == 4) *reinterpret_cast<uint32_t*>(a.sym + orow) = symw[0];
                else *reinterpret_cast<uint16_t*>(a.sym + orow) = (uint16_t)symw[0];
            }
            if (a.bits) {
                if (R == 8) {
                    reinterpret_cast<uint2*>(a.bits + 2 * orow)[0] = make_uint2(bitw[0], bitw[1]);
                    reinterpret_cast<uint2*>(a.bits + 2 * orow)[1] = make_uint2(bitw[2], bitw[3]);
                }
                else if (R == 4) *reinterpret_cast<uint2*>(a.bits + 2 * orow) = make_uint2(bitw[0], bitw[1]);
                else *reinterpret_cast<uint32_t*>(a.bits + 2 * orow) = bitw[0];
            }
            if (a.ref_bits) {
                uint32_t nerr = 0;
                if (ref_vec) { /* bit bytes against bit bytes, four at a time (only bit 0 of a reference byte counts, as in pack_symbol) */
#pragma unroll
                    for (int j = 0; j < (R + 1) / 2; ++j) nerr += __popc((bitw[j] ^ refw[j]) & 0x01010101u);
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const uint32_t h = bitw[r / 2] >> (16 * (r % 2));
                        nerr += __popc(pack_symbol(refp + 2 * r, 2) ^ (((h & 1u) << 1) | ((h >> 8) & 1u)));
                    }
                }
                err += nerr;
                cmp += 2 * R;
            }
        } else if (have_any) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (ka + r < a.K) {
                    const float I = __fmul_rn(a.rx_gain, ai[r]), Q = __fmul_rn(a.rx_gain, aq[r]);
                    const uint32_t s = slice_point4(s_slut + toff[r], a.n_const, I, Q);
                    err += emit_symbol(a, f, ka + r, s, I, Q);
                    cmp += a.ref_bits ? a.bps : 0u;
                }
            }
        }
        if (refp) refp += a.ref_stride;
        if (TXF && f + 1 < f1) park_bits(); /* read again only behind the barrier at the top of the next frame */
    }
    block_count(a, err, cmp);
    if (TMC > 0) tmem_free<(TMC > 0 ? TMC : 32)>(taddr);
}

/* ------------------------------------------------------------------ host side */
template <int NT, int OFF, bool FMA, bool NOISE, int THREADS, int MINB, int R, int PF = RX_DEFAULT_PF, int TMC = RX_DEFAULT_TMC, bool TXF = false>
cudaError_t rx_fast_launch_t(const RxArgs& a, const float* h_taps, cudaStream_t stream)
{
    using C = RxFastCfg<NT, OFF, THREADS, R>;
    const unsigned tiles = (unsigned)((a.K + C::TS - 1) / C::TS), groups = (unsigned)((a.F + a.frames_per_block - 1) / a.frames_per_block);
    dim3 grid = a.tile_major ? dim3(groups, tiles) : dim3(tiles, groups);
    const TapsParam<NT> tp = make_taps_param<NT>(h_taps);
    const size_t smem = C::smem(a.n_tables * a.n_const);
    auto kern = rx_fast_kernel<NT, OFF, FMA, NOISE, THREADS, MINB, R, PF, TMC, TXF>;
    /* the attributes are per DEVICE (and per instantiation): remember the shared-memory size each device was configured for */
    static std::atomic<size_t> configured[kMaxDevices];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kMaxDevices || configured[dev].load(std::memory_order_acquire) != smem + 1) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
#ifndef RX_CARVEOUT
#define RX_CARVEOUT 66 /* per cent of the SM's 228 KB: the 164 KB configuration holds 8 CTAs of the 64-thread shapes */
#endif
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, PF == 5 ? (int)(RX_CARVEOUT) : (int)cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < kMaxDevices) configured[dev].store(smem + 1, std::memory_order_release);
    }
    kern<<<grid, THREADS, smem, stream>>>(a, tp);
    return cudaGetLastError();
}

/* The instantiations of one (NT, THREADS, MINB, R, PF, TMC) shape, split into the noise-free and the noisy family so that
 * each lives in its own translation unit (the 129-tap noisy kernels take minutes to compile).
 * OFF shifts the tile so that it starts on an even sample (16-byte pairs): 0 for an odd decision delay, 1 for an even
 * one.  The noisy kernels prefer a tile that starts on a multiple of FOUR samples (one Philox block = one aligned quad =
 * the chunks of an even/odd lane pair, phase A); OFF = 3 provides that for decision delays that are multiples of 4 (the
 * matched-filter pair: delay 128), OFF = 0 already does for delay = 3 mod 4 (35); the other two residues run the
 * OFF = 0 / 1 kernels with every lane generating its own block.  tmem = false (MODEM_FLAG_NO_TMEM) selects instantiations
 * that allocate no tensor memory; they exist for the exact MAC only (rx_fast_supported). */
template <int NT, int THREADS, int MINB, int R, int PF, int TMC>
cudaError_t rx_fast_dispatch_clean(const RxArgs& a, const float* h_taps, bool fma, bool tmem, cudaStream_t stream)
{
    const bool odd = (a.delay & 1u) != 0;
#define MG_RX_CASE(O, F, T) return rx_fast_launch_t<NT, O, F, false, THREADS, MINB, R, PF, T>(a, h_taps, stream)
    if (!tmem) { if (odd) MG_RX_CASE(0, false, 0); else MG_RX_CASE(1, false, 0); }
    if (odd) { if (fma) MG_RX_CASE(0, true, TMC); else MG_RX_CASE(0, false, TMC); }
    if (fma) MG_RX_CASE(1, true, TMC); else MG_RX_CASE(1, false, TMC);
#undef MG_RX_CASE
}
template <int NT, int THREADS, int MINB, int R, int PF, int TMC>
cudaError_t rx_fast_dispatch_noise(const RxArgs& a, const float* h_taps, bool fma, bool tmem, cudaStream_t stream)
{
    const uint32_t off = (a.delay % 4u == 0) ? 3u : ((a.delay & 1u) ? 0u : 1u);
#define MG_RX_CASE(O, F, T) return rx_fast_launch_t<NT, O, F, true, THREADS, MINB, R, PF, T>(a, h_taps, stream)
#define MG_RX_OFF(O)                                                    \
    if (!tmem) MG_RX_CASE(O, false, 0);                                 \
    if (fma) MG_RX_CASE(O, true, TMC); else MG_RX_CASE(O, false, TMC);
    switch (off) {
    case 0: MG_RX_OFF(0)
    case 1: MG_RX_OFF(1)
    default: MG_RX_OFF(3)
    }
#undef MG_RX_OFF
#undef MG_RX_CASE
}

} /* namespace mg */
