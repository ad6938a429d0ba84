/*
 * modem_api.cu -- the C ABI of include/modem_gpu.h: context, buffer staging, kernel
 * selection and launch geometry.  No CPU fallback lives here: every compute entry needs
 * a usable sm_100 device and fails with an error code otherwise.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/modem_gpu.h"
#include "kernels.cuh"
#include "frontend.cuh"
#include "launch.h"

using mg::u64;

namespace {

struct Scratch {
    void* p = nullptr;
    size_t cap = 0;
};

} // namespace

struct modem_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    modem_cfg_t cfg{};
    std::vector<float> h_const, h_slut, h_tx_taps, h_rx_taps;
    float2* d_lut = nullptr;
    /* sign-product form of the shaped TX: every constellation point is (+-vi, +-vq) with one magnitude per rail, so
     * round(h[k] * (+-v)) = +-round(h[k] * v) is a per-tap constant and the reference's multiply-round-add-round of
     * fir.rs:21-24 becomes ONE fma(+-1, round(h[k] * v), acc) -- the product by +-1 is exact, so the single rounding
     * of the fma is the reference's rounding of the add: bit-identical with half the FP32 lane work */
    float2* d_sign_lut = nullptr;
    std::vector<float> h_tx_rail_taps; /* (round(h[k] * vi), round(h[k] * vq)) pairs */
    bool tx_sign_form = false;
    float2* d_slut = nullptr;
    float* d_tx_taps = nullptr;
    float* d_rx_taps = nullptr;
    float* d_chan_w = nullptr;
    float* d_chan_po = nullptr;
    size_t n_channels = 0, frames_per_channel = 1;
    u64* d_counters = nullptr;
    Scratch s_bits, s_tx, s_iq, s_rx, s_sym, s_bits_out, s_soft, s_filt;
    /* stateful phasor (modem_gpu_set_phasor) and the carrier-recovery front end */
    modem_phasor_t phasor{};
    bool phasor_on = false;
    Scratch s_state, s_re, s_raw, s_po, s_hilbert, s_siq, s_cptab;
    struct CpKey { /* what the cpfsk table was built for */
        u64 L = 0, sample0 = ~0ull;
        float deviation = 0.0f, amplitude = 0.0f;
        uint32_t bps = 0;
        bool operator==(const CpKey& o) const
        {
            return L == o.L && sample0 == o.sample0 && deviation == o.deviation && amplitude == o.amplitude && bps == o.bps;
        }
    } cp_key;
    /* host-buffer loopback pipeline: three streams by role (lanes[0] copies in, lanes[1] runs the kernels,
     * lanes[2] copies out) over a ring of chunk slots; the TX samples of a chunk never leave the kernel stream,
     * so one chunk-sized sample buffer serves every slot */
    struct Lane {
        cudaStream_t s = nullptr;
        cudaEvent_t done = nullptr;
    } lanes[3];
    std::vector<cudaEvent_t> pipe_events; /* 3 per chunk: copied in, computed, copied out (the last only when tracing) */
    cudaEvent_t pipe_t0 = nullptr;
    Scratch pipe_tx; /* one chunk of TX samples (two-kernel form only) */
    Scratch s_packed_in, s_packed_out; /* modem_gpu_loopback_packed: the call's packed rows on the device (host callers) */
    size_t packed_chunk = 0; /* MODEM_GPU_PACKED_CHUNK: frames per chunk of the packed-payload pipeline (0 = ~512 MB of samples) */
    bool lanes_ready = false;
    cudaEvent_t ev_start = nullptr;
    u64 frame_base = 0; /* see ChannelView::frame_base */
    /* NCO tables (ChannelView::cs_tab): TX view (no phase offset) and RX view (with it); cached by key */
    Scratch s_cs_tx, s_cs_rx;
    struct CsKey {
        u64 len = 0, ch0 = 0, nch = 0, ver = ~0ull, s0 = 0;
        bool operator==(const CsKey& o) const { return len == o.len && ch0 == o.ch0 && nch == o.nch && ver == o.ver && s0 == o.s0; }
    } cs_key;
    bool cs_rx_shared = true; /* every phase offset is 0: the RX view is the TX table */
    u64 chan_version = 0;
    /* bumped whenever something a captured loopback graph has baked in changes: the NCO tables (rebuilt or re-allocated),
     * the TX mapper (set_phasor), the stream (set_stream), the bank (set_channels) */
    u64 graph_epoch = 0;
    cudaEvent_t ev_pool[8] = {};
    size_t loop_chunk = 0; /* MODEM_GPU_LOOP_CHUNK: frames per chunk of the device loopback pipeline */
    /* the chunk pipeline of modem_gpu_loopback_device, captured once per argument set and replayed */
    struct LoopGraph {
        cudaGraphExec_t exec = nullptr;
        const void *bits = nullptr, *tx = nullptr, *sym = nullptr, *out = nullptr, *cnt = nullptr;
        size_t F = 0, nbits = 0, Fc = 0;
        float sigma = 0.0f;
        uint64_t seed = 0, frame0 = 0;
        u64 chan_version = 0, epoch = 0;
    } loop_graph;
    bool use_graph = true; /* MODEM_GPU_NO_GRAPH=1 disables */
    uint64_t launches = 0;
    bool force_generic = false;
    bool pipe_trace = false; /* MODEM_GPU_PIPE_TRACE=1: per-chunk event timeline of the host-buffer pipeline on stderr */
    bool pipe_fused = false; /* MODEM_GPU_PIPE_FUSED=1: the host-buffer pipeline runs the fused loopback kernel per chunk instead of TX + RX (measured slower, see loopback_pipelined) */
    int pipe_ramp = 0; /* MODEM_GPU_PIPE_RAMP: 1 = doubling chunks at both ends of the call, n >= 2 = one chunk of 1/n at each end (all measured slower or equal) */
    bool no_rx_dec = false; /* MODEM_GPU_NO_RX_DEC=1: shapes of the tuned any-sps RX kernel (rx_dec.cu) take the generic kernel */
    bool no_sign_slice = false; /* MODEM_GPU_NO_SIGN_SLICE=1: the fast RX kernel always runs the nearest-point search */
    bool no_fused_loop = false; /* MODEM_GPU_NO_FUSED_LOOP=1: the loopback entries run the TX and the RX kernel separately */
    int rx_fpb = 0, rx_tile_major = -1; /* MODEM_GPU_RX_FPB / MODEM_GPU_RX_TILEMAJOR: tuning knobs */
    size_t pipe_chunk = 0; /* MODEM_GPU_PIPE_CHUNK: frames per pipeline chunk (0 = ~64 MB of TX samples) */
    std::string last_error;
};

struct modem_comm {
    void* nccl_comm = nullptr;
    modem_ctx* ctx = nullptr;
    u64* d_buf = nullptr;
    size_t cap = 0;
};

namespace {

thread_local std::string g_last_error;

int fail(modem_ctx* ctx, int code, const std::string& msg)
{
    g_last_error = msg;
    if (ctx) ctx->last_error = msg;
    return code;
}

#define CK(ctx, call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail((ctx), MODEM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));    \
    } while (0)

bool is_device_ptr(const void* p)
{
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

int ensure(modem_ctx* ctx, Scratch& s, size_t bytes)
{
    if (bytes <= s.cap) return MODEM_OK;
    if (s.p) CK(ctx, cudaFree(s.p));
    s.p = nullptr;
    s.cap = 0;
    cudaError_t e = cudaMalloc(&s.p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, MODEM_ERR_NOMEM, "cudaMalloc of " + std::to_string(bytes) + " bytes failed");
    }
    s.cap = bytes;
    return MODEM_OK;
}

/* A buffer argument resolved to a device pointer: in place if the caller passed device
 * memory, else staged through context scratch. */
struct Staged {
    void* dev = nullptr;
    void* host = nullptr; /* non-null => copy back / from */
    size_t bytes = 0;
};

int stage_in(modem_ctx* ctx, Scratch& s, const void* p, size_t bytes, Staged* out)
{
    out->bytes = bytes;
    if (!p || bytes == 0) return MODEM_OK;
    if (is_device_ptr(p)) {
        out->dev = const_cast<void*>(p);
        return MODEM_OK;
    }
    int rc = ensure(ctx, s, bytes);
    if (rc) return rc;
    CK(ctx, cudaMemcpyAsync(s.p, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
    out->dev = s.p;
    return MODEM_OK;
}

int stage_out(modem_ctx* ctx, Scratch& s, void* p, size_t bytes, Staged* out)
{
    out->bytes = bytes;
    if (!p || bytes == 0) return MODEM_OK;
    if (is_device_ptr(p)) {
        out->dev = p;
        return MODEM_OK;
    }
    int rc = ensure(ctx, s, bytes);
    if (rc) return rc;
    out->dev = s.p;
    out->host = p;
    return MODEM_OK;
}

int finish_out(modem_ctx* ctx, const Staged& st)
{
    if (st.host && st.bytes) CK(ctx, cudaMemcpyAsync(st.host, st.dev, st.bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return MODEM_OK;
}

mg::ChannelView channel_view(const modem_ctx* ctx)
{
    mg::ChannelView v;
    v.w = ctx->n_channels ? ctx->d_chan_w : nullptr;
    v.po = ctx->n_channels ? ctx->d_chan_po : nullptr;
    v.w0 = ctx->cfg.sample_freq;
    v.po0 = ctx->cfg.phase_offset;
    v.frames_per_channel = ctx->n_channels ? ctx->frames_per_channel : 1;
    v.frame_base = ctx->frame_base;
    v.cs_tab = nullptr;
    v.cs_len = 0;
    v.cs_ch0 = 0;
    v.po_frame = nullptr;
    return v;
}

/* frames handled by one CTA: large enough to amortise the per-tile NCO table, small enough
 * to leave several waves of CTAs; must divide frames_per_channel so a CTA sees one carrier. */
uint32_t frames_per_block(const modem_ctx* ctx, u64 F, u64 tiles_x)
{
    const u64 target_ctas = (u64)ctx->sm_count * 16;
    u64 fpb = (F * tiles_x) / std::max<u64>(target_ctas, 1);
    fpb = std::min<u64>(std::max<u64>(fpb, std::min<u64>(F, 8)), 32);
    fpb = std::max<u64>(fpb, (F + 65534) / 65535); /* gridDim.y limit */
    if (ctx->n_channels) {
        u64 fc = ctx->frames_per_channel;
        if (fpb > fc) fpb = fc;
        while (fc % fpb || ctx->frame_base % fpb) --fpb; /* a CTA's frames must share one carrier */
    }
    return (uint32_t)fpb;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

/* Sign slicer of the fast RX kernel (rx_fast.cuh phase C).  Eligible when the gain-scaled slicer table is the four points
 * (-+A, -+B) in index order (bit 1 of the index = sign of i, bit 0 = sign of q: qpsk.rs:23-35 at phase 0), A and B
 * positive, finite and within a factor 2 of each other.  For soft values with |I|, |Q| in [min/1024, 4 min] the sign
 * test then returns exactly what the nearest-point search returns, roundings and tie rule included (the proof is in
 * DESIGN.md section 4); outside that window the kernel runs the search. */
void set_sign_slicer(const modem_ctx* ctx, mg::RxArgs& a);

/*
 * Make sure the NCO tables cover the channels of frames [frame_base, frame_base + F) at `len` samples
 * per frame, then point the view at the TX (rx == false) or RX table.  One tiny kernel per new key.
 */
constexpr u64 kCsPadLo = 160, kCsPadHi = 4352; /* zero margins so RX tiles can read past both frame ends unguarded */

int attach_carrier_table(modem_ctx* ctx, mg::ChannelView& view, u64 F, u64 len, bool rx, u64 sample_skip = 0)
{
    const u64 s0 = ctx->cfg.sample0 + sample_skip; /* Carrier.sample of the frames' first sample */
    const u64 fpc = ctx->n_channels ? ctx->frames_per_channel : 1;
    const u64 ch0 = ctx->n_channels ? ctx->frame_base / fpc : 0;
    const u64 ch1 = ctx->n_channels ? (ctx->frame_base + F - 1) / fpc : 0;
    modem_ctx::CsKey key;
    key.len = len;
    key.ver = ctx->chan_version;
    key.s0 = s0;
    /* keep a wider cached range if it already covers this call */
    if (ctx->cs_key.len == len && ctx->cs_key.ver == key.ver && ctx->cs_key.s0 == s0 && ctx->cs_key.ch0 <= ch0 &&
        ch1 < ctx->cs_key.ch0 + ctx->cs_key.nch) {
        key = ctx->cs_key;
    } else {
        key.ch0 = ch0;
        key.nch = ctx->n_channels ? std::min<u64>(ctx->n_channels - ch0, std::max<u64>(ch1 - ch0 + 1, 1)) : 1;
        const u64 row = kCsPadLo + len + kCsPadHi;
        const size_t bytes = key.nch * row * sizeof(float2);
        if (bytes > ((size_t)1 << 30)) return MODEM_OK; /* too many carriers for a table: kernels evaluate the NCO themselves */
        /* the cached key is void from here on: if anything below fails the tables hold no complete row set, and a captured
         * graph that baked in their addresses or contents must not be replayed */
        ctx->cs_key.ver = ~0ull;
        ctx->graph_epoch++;
        int rc = ensure(ctx, ctx->s_cs_tx, bytes);
        if (rc) return rc;
        mg::ChannelView cv = channel_view(ctx);
        const unsigned blocks = (unsigned)std::min<u64>((key.nch * row + mg::kThreads - 1) / mg::kThreads, (u64)ctx->sm_count * 16);
        mg::carrier_table_kernel<<<blocks, mg::kThreads, 0, ctx->stream>>>((float2*)ctx->s_cs_tx.p, len, kCsPadLo, kCsPadHi, key.ch0, key.nch, cv, s0, 0);
        ctx->launches++;
        if (!ctx->cs_rx_shared) {
            rc = ensure(ctx, ctx->s_cs_rx, bytes);
            if (rc) return rc;
            mg::carrier_table_kernel<<<blocks, mg::kThreads, 0, ctx->stream>>>((float2*)ctx->s_cs_rx.p, len, kCsPadLo, kCsPadHi, key.ch0, key.nch, cv, s0, 1);
            ctx->launches++;
        }
        CK(ctx, cudaGetLastError());
        ctx->cs_key = key;
    }
    view.cs_tab = (const float2*)((rx && !ctx->cs_rx_shared) ? ctx->s_cs_rx.p : ctx->s_cs_tx.p) + kCsPadLo; /* -> sample 0 of row 0 */
    view.cs_len = kCsPadLo + len + kCsPadHi;
    view.cs_ch0 = key.ch0;
    return MODEM_OK;
}

void set_sign_slicer(const modem_ctx* ctx, mg::RxArgs& a)
{
    a.sign_slice = 0;
    a.ss_lo = a.ss_hi = 0.0f;
    if (ctx->no_sign_slice || ctx->cfg.bits_per_symbol != 2 || ctx->cfg.n_tables != 1 || ctx->h_slut.size() != 8) return;
    const float A = ctx->h_slut[6], B = ctx->h_slut[7]; /* entry 3 = (+A, +B) */
    if (!(A > 0.0f) || !(B > 0.0f) || !std::isfinite(A) || !std::isfinite(B)) return;
    const float mn = std::min(A, B), mx = std::max(A, B);
    if (mx > 2.0f * mn || mn < 1e-30f || mx > 1e30f) return;
    for (int j = 0; j < 4; ++j) {
        const float wi = (j & 2) ? A : -A, wq = (j & 1) ? B : -B;
        if (memcmp(&wi, &ctx->h_slut[2 * j], 4) || memcmp(&wq, &ctx->h_slut[2 * j + 1], 4)) return;
    }
    a.sign_slice = 1;
    a.ss_lo = mn * (1.0f / 1024.0f);
    a.ss_hi = 4.0f * mn;
}

/* ------------------------------------------------------------------ TX launch */
/* the real-part wire output of src/bin/modulate.rs (generic kernels only) */
struct RealOut {
    float* re = nullptr;
    u64 stride = 0, offset = 0;
    u64 sample_skip = 0; /* samples the shared Carrier already produced (the preamble, modulate.rs:120,128) */
};

int launch_tx_phasor(modem_ctx* ctx, mg::TxArgs& a);

int launch_tx(modem_ctx* ctx, const uint8_t* d_bits, u64 F, u64 nbits, float2* d_tx, float2* d_iq, const RealOut* ro = nullptr)
{
    const modem_cfg_t& c = ctx->cfg;
    mg::TxArgs a{};
    if (ro) {
        a.re = ro->re;
        a.re_stride = ro->stride;
        a.re_offset = ro->offset;
    }
    a.bits = d_bits;
    a.nbits = nbits;
    a.tx = d_tx;
    a.iq = d_iq;
    a.nsym = nbits / c.bits_per_symbol;
    a.L = a.nsym * c.samples_per_symbol;
    a.F = F;
    a.lut = ctx->d_lut;
    a.bps = c.bits_per_symbol;
    a.sps = c.samples_per_symbol;
    a.n_tables = c.n_tables;
    a.n_const = 1u << c.bits_per_symbol;
    a.q_offset = c.q_offset;
    a.ch = channel_view(ctx);
    a.sample0 = c.sample0 + (ro ? ro->sample_skip : 0);
    a.taps = ctx->d_tx_taps;
    a.n_taps = c.n_tx_taps;
    if (F == 0 || a.L == 0) return MODEM_OK;
    if (a.nsym >= (1ull << 32)) return fail(ctx, MODEM_ERR_UNSUPPORTED, "more than 2^32 symbols per frame");
    if (ctx->phasor_on) return launch_tx_phasor(ctx, a);
    const bool fma = (c.flags & MODEM_FLAG_FUSED_MAC) != 0;
    const bool vec_ok = (a.L % 2 == 0) && aligned16(d_tx) && aligned16(d_iq);
    const bool plain = ro == nullptr; /* the tuned kernels write complex samples and take the carrier from the NCO table */

    if (c.n_tx_taps == 0) {
        const uint32_t bps = c.bits_per_symbol;
        const bool word_ok = mg::tx_rect_fast_supported(bps) && (nbits % bps == 0) &&
                             ((reinterpret_cast<uintptr_t>(d_bits) % bps) == 0);
        const bool fast_shape = !ctx->force_generic && !d_iq && c.q_offset == 0 && c.n_tables == 1 &&
                                word_ok && a.L < (1ull << 32) && (a.L % 2 == 0);
        const bool fast = fast_shape && ((plain && vec_ok && d_tx) || (ro && ro->re && !d_tx));
        if (fast) {
            a.frames_per_block = frames_per_block(ctx, F, mg::tx_rect_fast_tiles(a.L));
            int rc = attach_carrier_table(ctx, a.ch, F, a.L, false, ro ? ro->sample_skip : 0);
            if (rc) return rc;
            CK(ctx, mg::tx_rect_fast_launch(a, ctx->stream));
        } else {
            const int vec = vec_ok ? 2 : 1;
            const u64 tile = (u64)mg::kThreads * 2 * vec;
            const u64 tiles = (a.L + tile - 1) / tile;
            a.frames_per_block = frames_per_block(ctx, F, tiles);
            dim3 grid((unsigned)tiles, (unsigned)((F + a.frames_per_block - 1) / a.frames_per_block));
            if (vec == 2) mg::tx_rect_kernel<2><<<grid, mg::kThreads, 0, ctx->stream>>>(a);
            else mg::tx_rect_kernel<1><<<grid, mg::kThreads, 0, ctx->stream>>>(a);
        }
    } else if (plain && !ctx->force_generic && mg::tx_shaped_fast_supported(c.samples_per_symbol, c.n_tx_taps) && c.q_offset == 0 &&
               d_tx && !d_iq && vec_ok) {
        a.frames_per_block = frames_per_block(ctx, F, mg::tx_shaped_fast_tiles(a.nsym));
        int rc = attach_carrier_table(ctx, a.ch, F, a.L, false);
        if (rc) return rc;
        if (ctx->tx_sign_form && !fma) { /* exact, one FFMA2 per tap pair instead of FMUL2 + FFMA2 */
            a.lut = ctx->d_sign_lut;
            CK(ctx, mg::tx_shaped_fast_launch(a, ctx->h_tx_rail_taps.data(), true, true, ctx->stream));
        } else {
            CK(ctx, mg::tx_shaped_fast_launch(a, ctx->h_tx_taps.data(), fma, false, ctx->stream));
        }
    } else {
        const uint32_t sps = c.samples_per_symbol, N = c.n_tx_taps;
        uint32_t TS = std::max<uint32_t>(1, std::min<uint32_t>(256, 2048 / sps));
        a.sym_tile = TS;
        const uint32_t H = (N + sps - 1) / sps + 1;
        const size_t smem = (size_t)TS * sps * sizeof(float2) + (size_t)N * 4 + 2 * (size_t)(TS + H) * 4;
        if (smem > 200 * 1024) return fail(ctx, MODEM_ERR_UNSUPPORTED, "tx_taps too long for shared memory");
        const u64 tiles = (a.nsym + TS - 1) / TS;
        a.frames_per_block = frames_per_block(ctx, F, tiles);
        dim3 grid((unsigned)tiles, (unsigned)((F + a.frames_per_block - 1) / a.frames_per_block));
        if (fma) {
            CK(ctx, cudaFuncSetAttribute(mg::tx_shaped_generic_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            mg::tx_shaped_generic_kernel<true><<<grid, mg::kThreads, smem, ctx->stream>>>(a);
        } else {
            CK(ctx, cudaFuncSetAttribute(mg::tx_shaped_generic_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            mg::tx_shaped_generic_kernel<false><<<grid, mg::kThreads, smem, ctx->stream>>>(a);
        }
    }
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    return MODEM_OK;
}

/* ------------------------------------------------------------------ RX launch */
/* real-valued input of the demodulate binary (generic kernels only): see RxArgs::raw */
struct RawSrc {
    const void* raw = nullptr;
    uint32_t fmt = 0;
    u64 stride = 0, skip = 0;
    const float* d_po = nullptr; /* per-frame PLL.phase_offset (nullable) */
};

int launch_rx(modem_ctx* ctx, const float2* d_rx, u64 F, u64 L, uint8_t* d_sym, uint8_t* d_bits, float2* d_soft,
              float2* d_filt, const uint8_t* d_ref, u64 ref_stride, u64* d_counters, float sigma, uint64_t seed,
              uint64_t frame0, const RawSrc* src = nullptr)
{
    const modem_cfg_t& c = ctx->cfg;
    mg::RxArgs a{};
    a.rx = d_rx;
    a.L = L;
    a.F = F;
    a.K = modem_gpu_decided_symbols(ctx, L);
    a.sym = d_sym;
    a.bits = d_bits;
    a.soft = d_soft;
    a.filt = d_filt;
    a.ref_bits = d_ref;
    a.ref_stride = ref_stride;
    a.counters = d_counters;
    a.slut = ctx->d_slut;
    a.bps = c.bits_per_symbol;
    a.sps = c.samples_per_symbol;
    a.n_tables = c.n_tables;
    a.n_const = 1u << c.bits_per_symbol;
    a.q_offset = c.q_offset;
    a.delay = c.decision_delay;
    a.rx_gain = c.rx_gain;
    a.ch = channel_view(ctx);
    a.sample0 = c.sample0;
    a.taps = ctx->d_rx_taps;
    a.n_taps = c.n_rx_taps;
    a.nz = mg::make_noise(sigma, seed, frame0);
    if (src) {
        a.raw = src->raw;
        a.rx_fmt = src->fmt;
        a.raw_stride = src->stride;
        a.raw_skip = src->skip;
        a.sample0 = c.sample0 + src->skip; /* the demodulator's Carrier ran through the lock (demodulator.rs:34) */
        a.ch.po_frame = src->d_po;
    }
    if (F == 0 || L == 0) return MODEM_OK;
    const bool fma = (c.flags & MODEM_FLAG_FUSED_MAC) != 0;
    const uint32_t N = c.n_rx_taps, sps = c.samples_per_symbol;

    if (d_filt && !ctx->force_generic && mg::rx_fullrate_fast_supported(N) && (L % 2 == 0) && aligned16(d_filt) && a.nz.sigma == 0.0f) {
        /* the tuned full-rate kernel: sliding register window, packed MACs (rx_fullrate_fast.cu) */
        const bool pfp = a.ch.po_frame != nullptr;
        a.frames_per_block = std::min<uint32_t>(frames_per_block(ctx, F, mg::rx_fullrate_fast_tiles(L)), 8);
        if (ctx->n_channels) while (ctx->frames_per_channel % a.frames_per_block || ctx->frame_base % a.frames_per_block) --a.frames_per_block;
        if (!pfp && !src) { /* per-call phase offset: NCO from the context's table (one slice per channel) */
            int rc = attach_carrier_table(ctx, a.ch, F, L, true);
            if (rc) return rc;
        }
        CK(ctx, mg::rx_fullrate_fast_launch(a, ctx->h_rx_taps.data(), fma, pfp, ctx->stream));
        ctx->launches++;
    } else if (d_filt) {
        const uint32_t TILE = mg::kThreads * 4;
        const size_t smem = (size_t)(TILE + N - 1) * 8 + (size_t)N * 4;
        if (smem > 200 * 1024) return fail(ctx, MODEM_ERR_UNSUPPORTED, "rx_taps too long for shared memory");
        dim3 grid((unsigned)((L + TILE - 1) / TILE), (unsigned)F);
        if (F > 65535) return fail(ctx, MODEM_ERR_UNSUPPORTED, "full-rate dump limited to 65535 frames per call");
        if (fma) {
            CK(ctx, cudaFuncSetAttribute(mg::rx_fullrate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            mg::rx_fullrate_kernel<true><<<grid, mg::kThreads, smem, ctx->stream>>>(a);
        } else {
            CK(ctx, cudaFuncSetAttribute(mg::rx_fullrate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            mg::rx_fullrate_kernel<false><<<grid, mg::kThreads, smem, ctx->stream>>>(a);
        }
        ctx->launches++;
        CK(ctx, cudaGetLastError());
    }
    if (a.K == 0 || !(d_sym || d_bits || d_soft || d_ref)) return MODEM_OK;

    const bool fast_ok = !src && !ctx->force_generic && sps == 8 && c.q_offset == 0 && (L % 2 == 0) && aligned16(d_rx) &&
                         mg::rx_fast_supported(N, fma, !(c.flags & MODEM_FLAG_NO_TMEM));
    /* the real-valued wire formats of the demodulate binary (f32 / i16 rows, lock samples skipped, one phase offset per
     * frame): the fast sps-8 kernel with the NCO evaluated per frame (rx_fast_raw.cu).  Without a lock (one offset for the
     * whole call) the generic kernel, which evaluates the NCO once per CTA, is the faster one: measured */
    const bool raw_ok = src && src->d_po && !ctx->force_generic && sps == 8 && c.q_offset == 0 && !fma && sigma == 0.0f &&
                        !(c.flags & MODEM_FLAG_NO_TMEM) && mg::rx_fast_raw_supported(N, src->fmt) && L < (1ull << 31);
    if (raw_ok) {
        a.frames_per_block = std::min<uint32_t>(frames_per_block(ctx, F, mg::rx_fast_raw_tiles(a.K)), 16);
        if (ctx->n_channels) while (ctx->frames_per_channel % a.frames_per_block || ctx->frame_base % a.frames_per_block) --a.frames_per_block;
        a.tile_major = ctx->rx_tile_major == 0 ? 0u : 1u;
        set_sign_slicer(ctx, a);
        CK(ctx, mg::rx_fast_raw_launch(a, ctx->h_rx_taps.data(), ctx->stream));
    } else if (src && src->d_po && !ctx->force_generic && !ctx->no_rx_dec && sps != 8 && c.q_offset == 0 && !fma && sigma == 0.0f &&
               !(c.flags & MODEM_FLAG_NO_TMEM) && mg::rx_dec_raw_supported(N, sps, src->fmt) && L < (1ull << 31)) {
        /* the same wire at any other samples-per-symbol count (the reference's own rates: 45): rx_dec_kernel<..., RAW> */
        a.sym_tile = mg::rx_dec_tile_symbols(sps);
        a.frames_per_block = frames_per_block(ctx, F, (a.K + a.sym_tile - 1) / a.sym_tile);
        set_sign_slicer(ctx, a);
        CK(ctx, mg::rx_dec_raw_launch(a, ctx->h_rx_taps.data(), ctx->stream));
    } else if (fast_ok) {
        a.frames_per_block = frames_per_block(ctx, F, mg::rx_fast_tiles(N, a.K));
        if (ctx->rx_fpb > 0 && !ctx->n_channels) a.frames_per_block = (uint32_t)std::max<u64>(ctx->rx_fpb, (F + 65534) / 65535);
        else a.frames_per_block = std::min<uint32_t>(a.frames_per_block, 16); /* measured: 8..16 is the sweet spot once the NCO table removed the per-CTA setup */
        if (ctx->n_channels) while (ctx->frames_per_channel % a.frames_per_block || ctx->frame_base % a.frames_per_block) --a.frames_per_block;
        a.tile_major = ctx->rx_tile_major == 0 ? 0u : 1u; /* default on: co-resident CTAs share one NCO table slice in L1 */
        int rc = attach_carrier_table(ctx, a.ch, F, L, true);
        if (rc) return rc;
        if (!a.ch.cs_tab) return fail(ctx, MODEM_ERR_UNSUPPORTED, "carrier bank too large for the NCO table (> 1 GiB)");
        set_sign_slicer(ctx, a);
        CK(ctx, mg::rx_fast_launch(a, ctx->h_rx_taps.data(), fma, !(c.flags & MODEM_FLAG_NO_TMEM), ctx->stream));
    } else if (!src && !ctx->force_generic && !ctx->no_rx_dec && c.q_offset == 0 && sigma == 0.0f && (L % 2 == 0) && aligned16(d_rx) &&
               mg::rx_dec_supported(N, sps)) {
        /* any other samples-per-symbol count with the 64-tap low-pass (the reference's default rates: sps 45) */
        a.sym_tile = mg::rx_dec_tile_symbols(sps);
        a.frames_per_block = frames_per_block(ctx, F, (a.K + a.sym_tile - 1) / a.sym_tile);
        int rc = attach_carrier_table(ctx, a.ch, F, L, true);
        if (rc) return rc;
        if (!a.ch.cs_tab) return fail(ctx, MODEM_ERR_UNSUPPORTED, "carrier bank too large for the NCO table (> 1 GiB)");
        set_sign_slicer(ctx, a);
        CK(ctx, mg::rx_dec_launch(a, ctx->h_rx_taps.data(), fma, !(c.flags & MODEM_FLAG_NO_TMEM), ctx->stream));
    } else {
        size_t budget = 96 * 1024;
        if ((size_t)N * 4 + (size_t)(N + c.q_offset) * 17 + 64 > budget)
            return fail(ctx, MODEM_ERR_UNSUPPORTED, "rx_taps too long for shared memory");
        /* up to 128 symbols x 2 rails keep all 256 threads busy in the FIR; tiles of <= 48 KB leave room for 4 CTAs
         * per SM so that one CTA's staging overlaps another's FIR (long filters may need more for a useful tile) */
        budget = std::min<size_t>(budget, std::max<size_t>(48 * 1024, (size_t)N * 4 + 64 + 17 * ((size_t)15 * sps + N + c.q_offset)));
        const size_t rmax = (budget - (size_t)N * 4 - 64) / 17; /* 8 B of NCO + 2 skewed rails (4 + 1/8 B each) per staged sample */
        uint32_t TS = (uint32_t)std::min<size_t>(128, (rmax - N - c.q_offset) / sps + 1);
        TS = std::max<uint32_t>(TS, 1);
        a.sym_tile = TS;
        const size_t R = (size_t)(TS - 1) * sps + N + c.q_offset;
        const size_t RP = mg::rx_generic_skew((uint32_t)R) + 1;
        const size_t smem = R * 8 + 2 * RP * 4 + (size_t)N * 4;
        const u64 tiles = (a.K + TS - 1) / TS;
        a.frames_per_block = frames_per_block(ctx, F, tiles);
        /* per-frame phase offsets (PLL lock): the CTA keeps the frame-invariant NCO phase of its tile and adds
         * each frame's offset before the sincos, so it still loops over several frames */
        dim3 grid((unsigned)tiles, (unsigned)((F + a.frames_per_block - 1) / a.frames_per_block));
        if (fma) {
            CK(ctx, cudaFuncSetAttribute(mg::rx_generic_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            mg::rx_generic_kernel<true><<<grid, mg::kThreads, smem, ctx->stream>>>(a);
        } else {
            CK(ctx, cudaFuncSetAttribute(mg::rx_generic_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            mg::rx_generic_kernel<false><<<grid, mg::kThreads, smem, ctx->stream>>>(a);
        }
    }
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    return MODEM_OK;
}


/* ------------------------------------------------------------------ stateful phasors (TX) */
/*
 * Fused loopback: one kernel makes the TX samples of a tile from the bits, stores them (the TX buffer the caller
 * gets is the same, bit for bit) and demodulates them from registers -- the RX side's 8 B/sample read is gone.
 * Eligible: the headline shape (any one 4-point table -- QPSK at any phase --, rectangular hold, sps 8, the 64-tap low-pass,
 * odd decision delay, exact MACs, no noise, no phase offset anywhere so that both sides share one NCO table).  Returns 1 = launched,
 * 0 = not eligible (the caller runs the two kernels), < 0 = error.
 */
/* the shape test of the fused loopback; fills the geometry of `a`.  d_tx may be null: the samples are then not stored
 * at all (a caller that did not ask for the TX buffer gets a loopback without any sample traffic) */
bool loop_fused_eligible(modem_ctx* ctx, const uint8_t* d_bits, u64 F, u64 nbits, const float2* d_tx, float sigma, mg::RxArgs& a)
{
    const modem_cfg_t& c = ctx->cfg;
    if (ctx->no_fused_loop || ctx->force_generic || ctx->phasor_on || sigma != 0.0f) return false;
    if (c.n_tx_taps || c.bits_per_symbol != 2 || c.n_tables != 1 || c.q_offset || c.n_rx_taps != 64 ||
        (c.flags & MODEM_FLAG_FUSED_MAC) || !ctx->cs_rx_shared)
        return false;
    a.sps = c.samples_per_symbol;
    a.L = (nbits / 2) * a.sps;
    a.F = F;
    a.K = modem_gpu_decided_symbols(ctx, a.L);
    if (F == 0 || a.L == 0 || a.K == 0 || a.L >= (1ull << 32)) return false;
    a.delay = c.decision_delay;
    if (a.sps == 8) { /* the headline shape: rx_fast_kernel<..., TXF>, rows of bits read as 8-byte words */
        if (!d_bits || (nbits & 7u) || (reinterpret_cast<uintptr_t>(d_bits) & 7u) || (d_tx && !aligned16(d_tx))) return false;
        if (mg::loop_fused_supported_64(a)) return true;
    }
    /* any other samples-per-symbol count up to 64 (the reference's default rates: 45): rx_dec_kernel<..., TXF>, two bit bytes
     * per symbol read as one 16-bit word */
    if (ctx->no_rx_dec || !d_bits || (nbits & 1u) || (reinterpret_cast<uintptr_t>(d_bits) & 1u) || (d_tx && !aligned16(d_tx))) return false;
    a.sym_tile = mg::rx_dec_tile_symbols(a.sps);
    return a.sps != 8 && mg::loop_fused_dec_supported(a);
}

int launch_loop_fused(modem_ctx* ctx, const uint8_t* d_bits, u64 F, u64 nbits, float2* d_tx, uint8_t* d_sym, uint8_t* d_out,
                      u64* d_counters, float sigma)
{
    const modem_cfg_t& c = ctx->cfg;
    mg::RxArgs a{};
    if (!d_counters || !loop_fused_eligible(ctx, d_bits, F, nbits, d_tx, sigma, a)) return 0;
    const bool dec = a.sps != 8;
    a.sym = d_sym;
    a.bits = d_out;
    a.ref_bits = d_bits;
    a.ref_stride = nbits;
    a.counters = d_counters;
    a.slut = ctx->d_slut;
    a.bps = 2;
    a.n_tables = 1;
    a.n_const = 4;
    a.delay = c.decision_delay;
    a.rx_gain = c.rx_gain;
    a.ch = channel_view(ctx);
    a.sample0 = c.sample0;
    a.taps = ctx->d_rx_taps;
    a.n_taps = 64;
    a.tx_out = d_tx;
    for (int j = 0; j < 4; ++j) a.tx_iq[j] = make_float2(ctx->h_const[2 * j], ctx->h_const[2 * j + 1]);
    if (dec) {
        a.frames_per_block = frames_per_block(ctx, F, (a.K + a.sym_tile - 1) / a.sym_tile);
    } else {
        const u64 tiles = (a.K + mg::loop_fused_tile_symbols_64() - 1) / mg::loop_fused_tile_symbols_64();
        a.frames_per_block = std::min<uint32_t>(frames_per_block(ctx, F, tiles), 16);
        if (ctx->rx_fpb > 0) a.frames_per_block = (uint32_t)std::max<u64>(ctx->rx_fpb, (F + 65534) / 65535);
        if (ctx->n_channels) while (ctx->frames_per_channel % a.frames_per_block || ctx->frame_base % a.frames_per_block) --a.frames_per_block;
        a.tile_major = ctx->rx_tile_major == 0 ? 0u : 1u;
    }
    int rc = attach_carrier_table(ctx, a.ch, F, a.L, true);
    if (rc) return rc;
    if (!a.ch.cs_tab) return 0;
    set_sign_slicer(ctx, a);
    const bool tmem = !(c.flags & MODEM_FLAG_NO_TMEM);
    cudaError_t e = dec ? mg::loop_fused_dec_launch(a, ctx->h_rx_taps.data(), tmem, ctx->stream)
                        : mg::loop_fused_launch_64(a, ctx->h_rx_taps.data(), tmem, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, MODEM_ERR_CUDA, std::string("fused loopback: ") + cudaGetErrorString(e));
    ctx->launches++;
    return 1;
}

int launch_tx_phasor(modem_ctx* ctx, mg::TxArgs& a)
{
    const modem_phasor_t& ph = ctx->phasor;
    mg::PhasorArgs p{};
    p.kind = ph.kind;
    p.bps = ph.bits_per_symbol;
    p.increase_map = ph.mfsk_increase_map;
    p.max_symbol = (int)((1u << ph.bits_per_symbol) - 1u);
    p.amplitude = ph.amplitude;
    p.deviation = ph.deviation;
    p.phase = ph.phase;
    p.shift = ph.shift;
    p.samples_per_bit = (float)(a.sps / 2); /* msk.rs:17 */
    if (a.n_taps) return fail(ctx, MODEM_ERR_UNSUPPORTED, "tx_taps are not combined with a stateful phasor");
    const bool scan = ph.kind == MODEM_PHASOR_BFSK || ph.kind == MODEM_PHASOR_MFSK || ph.kind == MODEM_PHASOR_DMPSK;
    if (scan) {
        int rc = ensure(ctx, ctx->s_state, a.F * a.nsym * sizeof(float));
        if (rc) return rc;
        p.state = (const float*)ctx->s_state.p;
        const unsigned blocks = (unsigned)((a.F + 31) / 32);
        const size_t rowb = (size_t)mg::kScanChunk * p.bps + 4;
        const size_t smem = 2 * 32 * rowb + 2 * 32 * (mg::kScanChunk + 1) * sizeof(float);
        const int aligned4 = (a.nbits % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.bits) & 3u) == 0);
        float* d_state = (float*)ctx->s_state.p;
#define MG_SCAN(K)                                                                                                          \
    do {                                                                                                                    \
        if (smem > 48 * 1024) CK(ctx, cudaFuncSetAttribute(mg::phasor_scan_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        mg::phasor_scan_kernel<K><<<blocks, 64, smem, ctx->stream>>>(a.bits, a.nbits, a.F, a.nsym, a.sps, a.sample0, p, d_state, aligned4); \
    } while (0)
        if (ph.kind == MODEM_PHASOR_BFSK) MG_SCAN(mg::kPhBfsk);
        else if (ph.kind == MODEM_PHASOR_MFSK) MG_SCAN(mg::kPhMfsk);
        else MG_SCAN(mg::kPhDmpsk);
#undef MG_SCAN
        ctx->launches++;
        CK(ctx, cudaGetLastError());
        if (ph.kind == MODEM_PHASOR_DMPSK) {
            rc = ensure(ctx, ctx->s_siq, a.F * a.nsym * sizeof(float2));
            if (rc) return rc;
            const u64 n = a.F * a.nsym;
            const unsigned sb = (unsigned)std::min<u64>((n + mg::kThreads - 1) / mg::kThreads, (u64)ctx->sm_count * 32);
            mg::phasor_symbol_iq_kernel<<<sb, mg::kThreads, 0, ctx->stream>>>(p.state, n, p.amplitude, (float2*)ctx->s_siq.p);
            ctx->launches++;
            CK(ctx, cudaGetLastError());
            p.siq = (const float2*)ctx->s_siq.p;
        }
    }
    if (ph.kind == MODEM_PHASOR_CPFSK) {
        /* per-(symbol value, sample) table, shared by every frame: worth it once the frames outnumber the rows */
        const u64 n_sym = 1ull << ph.bits_per_symbol;
        const size_t bytes = n_sym * a.L * sizeof(float2);
        if (a.F >= 2 * n_sym && bytes <= ((size_t)256 << 20) && !ctx->n_channels) {
            modem_ctx::CpKey key{a.L, a.sample0, ph.deviation, ph.amplitude, ph.bits_per_symbol};
            if (!(key == ctx->cp_key)) {
                int rc = ensure(ctx, ctx->s_cptab, bytes);
                if (rc) return rc;
                const unsigned tb = (unsigned)std::min<u64>((n_sym * a.L + mg::kThreads - 1) / mg::kThreads, (u64)ctx->sm_count * 32);
                mg::cpfsk_table_kernel<<<tb, mg::kThreads, 0, ctx->stream>>>((float2*)ctx->s_cptab.p, a.L, (uint32_t)n_sym, p, a.sample0);
                ctx->launches++;
                CK(ctx, cudaGetLastError());
                ctx->cp_key = key;
            }
            p.cp_tab = (const float2*)ctx->s_cptab.p;
        }
    }
    const bool vec = (a.L % 2 == 0) && aligned16(a.tx) && aligned16(a.iq);
    const u64 tile = (u64)mg::kThreads * 2 * (vec ? 2 : 1);
    const u64 tiles = (a.L + tile - 1) / tile;
    a.frames_per_block = frames_per_block(ctx, a.F, tiles);
    dim3 grid((unsigned)tiles, (unsigned)((a.F + a.frames_per_block - 1) / a.frames_per_block));
#define MG_PH_LAUNCH(K)                                                                                   \
    do {                                                                                                  \
        if (vec) mg::tx_phasor_kernel<K, 2><<<grid, mg::kThreads, 0, ctx->stream>>>(a, p);                \
        else mg::tx_phasor_kernel<K, 1><<<grid, mg::kThreads, 0, ctx->stream>>>(a, p);                    \
    } while (0)
    switch (ph.kind) {
    case MODEM_PHASOR_BFSK: MG_PH_LAUNCH(mg::kPhBfsk); break;
    case MODEM_PHASOR_MFSK: MG_PH_LAUNCH(mg::kPhMfsk); break;
    case MODEM_PHASOR_CPFSK: MG_PH_LAUNCH(mg::kPhCpfsk); break;
    case MODEM_PHASOR_MSK: MG_PH_LAUNCH(mg::kPhMsk); break;
    case MODEM_PHASOR_DMPSK: MG_PH_LAUNCH(mg::kPhDmpsk); break;
    default: return fail(ctx, MODEM_ERR_INVALID, "unknown phasor kind");
    }
#undef MG_PH_LAUNCH
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    return MODEM_OK;
}

/* ------------------------------------------------------------------ PLL lock launch */
size_t sample_bytes(uint32_t fmt) { return fmt == MODEM_SAMPLES_C32 ? 8 : fmt == MODEM_SAMPLES_F32 ? 4 : 2; }

int launch_lock(modem_ctx* ctx, const void* d_samples, uint32_t fmt, u64 F, u64 stride, const float* hilbert, size_t n_h,
                size_t lock, float* d_po)
{
    mg::LockArgs a{};
    a.samples = d_samples;
    a.fmt = fmt;
    a.F = F;
    a.stride = stride;
    a.lock = (uint32_t)lock;
    a.ch = channel_view(ctx);
    a.sample0 = ctx->cfg.sample0;
    a.po = d_po;
    if (fmt != MODEM_SAMPLES_C32) {
        size_t n = n_h;
        const float* h = hilbert;
        if (!h) h = modem_hilbert_taps(&n);
        if (n == 0 || n > 4096) return fail(ctx, MODEM_ERR_INVALID, "lock_phase: bad Hilbert tap count");
        int rc = ensure(ctx, ctx->s_hilbert, n * sizeof(float));
        if (rc) return rc;
        CK(ctx, cudaMemcpyAsync(ctx->s_hilbert.p, h, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        a.htaps = (const float*)ctx->s_hilbert.p;
        a.n_h = (uint32_t)n;
    }
    if (F == 0) return MODEM_OK;
    mg::lock_phase_kernel<<<(unsigned)((F + 63) / 64), 64, 0, ctx->stream>>>(a);
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    return MODEM_OK;
}

int upload(modem_ctx* ctx, void** dptr, const void* src, size_t bytes)
{
    if (*dptr) {
        CK(ctx, cudaFree(*dptr));
        *dptr = nullptr;
    }
    if (!bytes) return MODEM_OK;
    cudaError_t e = cudaMalloc(dptr, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, MODEM_ERR_NOMEM, "cudaMalloc failed");
    }
    CK(ctx, cudaMemcpy(*dptr, src, bytes, cudaMemcpyHostToDevice));
    return MODEM_OK;
}

} // namespace

extern "C" {

int modem_gpu_abi_version(void) { return MODEM_GPU_ABI_VERSION; }

const char* modem_gpu_strerror(int code)
{
    switch (code) {
    case MODEM_OK: return "ok";
    case MODEM_ERR_INVALID: return "invalid argument or configuration";
    case MODEM_ERR_CUDA: return "CUDA runtime error";
    case MODEM_ERR_NOMEM: return "out of device memory";
    case MODEM_ERR_UNSUPPORTED: return "unsupported configuration";
    case MODEM_ERR_NCCL: return "NCCL error";
    case MODEM_ERR_NO_DEVICE: return "no usable sm_100 CUDA device (this library has no CPU fallback)";
    default: return "unknown error";
    }
}

const char* modem_gpu_last_error(const modem_ctx_t* ctx) { return ctx ? ctx->last_error.c_str() : g_last_error.c_str(); }

uint64_t modem_gpu_launch_count(const modem_ctx_t* ctx) { return ctx ? ctx->launches : 0; }

int modem_gpu_device_count(int* n)
{
    if (!n) return MODEM_ERR_INVALID;
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *n = 0;
        return fail(nullptr, MODEM_ERR_NO_DEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    *n = c;
    return MODEM_OK;
}

int modem_gpu_create(modem_ctx_t** out, int device, const modem_cfg_t* cfg)
{
    if (!out || !cfg) return fail(nullptr, MODEM_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->struct_size != sizeof(modem_cfg_t)) return fail(nullptr, MODEM_ERR_INVALID, "modem_cfg_t size mismatch (ABI)");
    if (cfg->bits_per_symbol < 1 || cfg->bits_per_symbol > 8) return fail(nullptr, MODEM_ERR_INVALID, "bits_per_symbol must be 1..8");
    if (cfg->samples_per_symbol < 1) return fail(nullptr, MODEM_ERR_INVALID, "samples_per_symbol must be >= 1");
    if (cfg->n_tables < 1 || (cfg->n_tables << cfg->bits_per_symbol) > (uint32_t)mg::kMaxLut || !cfg->const_iq)
        return fail(nullptr, MODEM_ERR_INVALID, "bad constellation table");
    if (cfg->n_rx_taps < 1 || !cfg->rx_taps) return fail(nullptr, MODEM_ERR_INVALID, "rx_taps required");
    if (cfg->n_tx_taps && !cfg->tx_taps) return fail(nullptr, MODEM_ERR_INVALID, "tx_taps missing");
    if (cfg->q_offset && (cfg->bits_per_symbol != 2 || cfg->q_offset >= cfg->samples_per_symbol))
        return fail(nullptr, MODEM_ERR_INVALID, "q_offset needs bits_per_symbol == 2 and q_offset < sps (data.rs:91-92)");

    int ndev = 0;
    int rc = modem_gpu_device_count(&ndev);
    if (rc) return rc;
    if (device < 0 || device >= ndev) return fail(nullptr, MODEM_ERR_NO_DEVICE, "no such CUDA device");
    cudaDeviceProp prop;
    CK(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(nullptr, MODEM_ERR_NO_DEVICE,
                    std::string("device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) + ", kernels are built for sm_100a only");
    CK(nullptr, cudaSetDevice(device));

    modem_ctx* ctx = new modem_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cfg = *cfg;
    const size_t np = (size_t)cfg->n_tables << cfg->bits_per_symbol;
    ctx->h_const.assign(cfg->const_iq, cfg->const_iq + 2 * np);
    ctx->h_slut.resize(2 * np);
    for (size_t i = 0; i < 2 * np; ++i) {
        volatile float v = cfg->slicer_gain * ctx->h_const[i]; /* one rounded binary32 product, as the oracle's g * c */
        ctx->h_slut[i] = v;
    }
    ctx->h_rx_taps.assign(cfg->rx_taps, cfg->rx_taps + cfg->n_rx_taps);
    if (cfg->n_tx_taps) ctx->h_tx_taps.assign(cfg->tx_taps, cfg->tx_taps + cfg->n_tx_taps);
    std::vector<float> h_sign;
    if (cfg->n_tx_taps && np) { /* sign-product form: one magnitude per rail over every table? */
        auto mag = [](float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0x7fffffffu; return u; };
        const uint32_t mi = mag(ctx->h_const[0]), mq = mag(ctx->h_const[1]);
        bool ok = true;
        for (size_t i = 0; i < np && ok; ++i) ok = mag(ctx->h_const[2 * i]) == mi && mag(ctx->h_const[2 * i + 1]) == mq;
        const char* ns = getenv("MODEM_GPU_NO_SIGN_FORM");
        if (ok && !(ns && ns[0] == '1')) {
            float vi, vq;
            memcpy(&vi, &mi, 4);
            memcpy(&vq, &mq, 4);
            h_sign.resize(2 * np);
            for (size_t i = 0; i < 2 * np; ++i) h_sign[i] = std::signbit(ctx->h_const[i]) ? -1.0f : 1.0f;
            ctx->h_tx_rail_taps.resize(2 * (size_t)cfg->n_tx_taps);
            for (uint32_t k = 0; k < cfg->n_tx_taps; ++k) {
                volatile float pi = ctx->h_tx_taps[k] * vi, pq = ctx->h_tx_taps[k] * vq; /* one rounded binary32 product each */
                ctx->h_tx_rail_taps[2 * k] = pi;
                ctx->h_tx_rail_taps[2 * k + 1] = pq;
            }
            ctx->tx_sign_form = true;
        }
    }
    ctx->cfg.const_iq = ctx->h_const.data();
    ctx->cfg.rx_taps = ctx->h_rx_taps.data();
    ctx->cfg.tx_taps = cfg->n_tx_taps ? ctx->h_tx_taps.data() : nullptr;
    const char* fg = getenv("MODEM_GPU_FORCE_GENERIC");
    ctx->force_generic = fg && fg[0] == '1';
    const char* pf = getenv("MODEM_GPU_PIPE_FUSED");
    ctx->pipe_fused = pf && pf[0] == '1';
    const char* pr = getenv("MODEM_GPU_PIPE_RAMP");
    ctx->pipe_ramp = pr ? atoi(pr) : 0;
    const char* ptr = getenv("MODEM_GPU_PIPE_TRACE");
    ctx->pipe_trace = ptr && ptr[0] == '1';
    const char* nf = getenv("MODEM_GPU_NO_FUSED_LOOP");
    ctx->no_fused_loop = nf && nf[0] == '1';
    const char* nrd = getenv("MODEM_GPU_NO_RX_DEC");
    ctx->no_rx_dec = nrd && nrd[0] == '1';
    const char* nss = getenv("MODEM_GPU_NO_SIGN_SLICE");
    ctx->no_sign_slice = nss && nss[0] == '1';
    const char* rf = getenv("MODEM_GPU_RX_FPB");
    ctx->rx_fpb = rf ? atoi(rf) : 0;
    const char* tm = getenv("MODEM_GPU_RX_TILEMAJOR");
    ctx->rx_tile_major = tm ? atoi(tm) : -1;
    const char* pc = getenv("MODEM_GPU_PIPE_CHUNK");
    ctx->pipe_chunk = pc ? (size_t)atoll(pc) : 0;
    const char* pkc = getenv("MODEM_GPU_PACKED_CHUNK");
    if (pkc && atoll(pkc) > 0) ctx->packed_chunk = (size_t)atoll(pkc);
    const char* lc = getenv("MODEM_GPU_LOOP_CHUNK");
    ctx->loop_chunk = lc ? (size_t)atoll(lc) : 0;
    const char* ng = getenv("MODEM_GPU_NO_GRAPH");
    ctx->use_graph = !(ng && ng[0] == '1');
    ctx->cs_rx_shared = cfg->phase_offset == 0.0f;

    rc = MODEM_OK;
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) rc = fail(nullptr, MODEM_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
    ctx->own_stream = rc == MODEM_OK;
    if (!rc) rc = upload(ctx, (void**)&ctx->d_lut, ctx->h_const.data(), ctx->h_const.size() * 4);
    if (!rc) rc = upload(ctx, (void**)&ctx->d_slut, ctx->h_slut.data(), ctx->h_slut.size() * 4);
    if (!rc && ctx->tx_sign_form) rc = upload(ctx, (void**)&ctx->d_sign_lut, h_sign.data(), h_sign.size() * 4);
    if (!rc) rc = upload(ctx, (void**)&ctx->d_rx_taps, ctx->h_rx_taps.data(), ctx->h_rx_taps.size() * 4);
    if (!rc && cfg->n_tx_taps) rc = upload(ctx, (void**)&ctx->d_tx_taps, ctx->h_tx_taps.data(), ctx->h_tx_taps.size() * 4);
    if (!rc) {
        e = cudaMalloc((void**)&ctx->d_counters, 2 * sizeof(u64));
        if (e != cudaSuccess) rc = fail(nullptr, MODEM_ERR_NOMEM, "cudaMalloc counters");
    }
    if (rc) {
        std::string msg = g_last_error;
        modem_gpu_destroy(ctx);
        g_last_error = msg;
        return rc;
    }
    *out = ctx;
    return MODEM_OK;
}

void modem_gpu_destroy(modem_ctx_t* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    void* ptrs[] = {ctx->d_lut, ctx->d_sign_lut, ctx->d_slut, ctx->d_tx_taps, ctx->d_rx_taps, ctx->d_chan_w, ctx->d_chan_po, ctx->d_counters,
                    ctx->s_bits.p, ctx->s_tx.p, ctx->s_iq.p, ctx->s_rx.p, ctx->s_sym.p, ctx->s_bits_out.p, ctx->s_soft.p, ctx->s_filt.p,
                    ctx->s_state.p, ctx->s_re.p, ctx->s_raw.p, ctx->s_po.p, ctx->s_hilbert.p, ctx->s_siq.p, ctx->s_cptab.p};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    for (auto& ln : ctx->lanes) {
        if (ln.s) cudaStreamSynchronize(ln.s);
        if (ln.done) cudaEventDestroy(ln.done);
        if (ln.s) cudaStreamDestroy(ln.s);
    }
    for (cudaEvent_t e : ctx->pipe_events)
        if (e) cudaEventDestroy(e);
    if (ctx->pipe_t0) cudaEventDestroy(ctx->pipe_t0);
    if (ctx->pipe_tx.p) cudaFree(ctx->pipe_tx.p);
    if (ctx->s_packed_in.p) cudaFree(ctx->s_packed_in.p);
    if (ctx->s_packed_out.p) cudaFree(ctx->s_packed_out.p);
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    for (auto e : ctx->ev_pool)
        if (e) cudaEventDestroy(e);
    if (ctx->loop_graph.exec) cudaGraphExecDestroy(ctx->loop_graph.exec);
    if (ctx->s_cs_tx.p) cudaFree(ctx->s_cs_tx.p);
    if (ctx->s_cs_rx.p) cudaFree(ctx->s_cs_rx.p);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    cudaGetLastError();
    delete ctx;
}

int modem_gpu_set_stream(modem_ctx_t* ctx, void* cuda_stream)
{
    if (!ctx) return MODEM_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    if (ctx->own_stream && ctx->stream) {
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        CK(ctx, cudaStreamDestroy(ctx->stream));
    }
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    ctx->own_stream = false;
    ctx->graph_epoch++;
    return MODEM_OK;
}

int modem_gpu_set_channels(modem_ctx_t* ctx, size_t n_channels, const float* sample_freq, const float* phase_offset,
                           size_t frames_per_channel)
{
    if (!ctx) return MODEM_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->graph_epoch++;
    if (n_channels == 0) {
        ctx->n_channels = 0;
        ctx->frames_per_channel = 1;
        ctx->chan_version++;
        ctx->cs_rx_shared = ctx->cfg.phase_offset == 0.0f;
        return MODEM_OK;
    }
    if (!sample_freq || frames_per_channel == 0) return fail(ctx, MODEM_ERR_INVALID, "set_channels: bad arguments");
    std::vector<float> po(n_channels, ctx->cfg.phase_offset);
    if (phase_offset) po.assign(phase_offset, phase_offset + n_channels);
    int rc = upload(ctx, (void**)&ctx->d_chan_w, sample_freq, n_channels * 4);
    if (!rc) rc = upload(ctx, (void**)&ctx->d_chan_po, po.data(), n_channels * 4);
    if (rc) return rc;
    ctx->n_channels = n_channels;
    ctx->frames_per_channel = frames_per_channel;
    ctx->chan_version++;
    ctx->cs_rx_shared = true;
    for (float v : po)
        if (v != 0.0f) ctx->cs_rx_shared = false;
    return MODEM_OK;
}

int modem_gpu_synchronize(modem_ctx_t* ctx)
{
    if (!ctx) return MODEM_ERR_INVALID;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return MODEM_OK;
}

size_t modem_gpu_frame_samples(const modem_ctx_t* ctx, size_t nbits)
{
    if (!ctx) return 0;
    return (nbits / ctx->cfg.bits_per_symbol) * ctx->cfg.samples_per_symbol;
}

size_t modem_gpu_decided_symbols(const modem_ctx_t* ctx, size_t L)
{
    if (!ctx) return 0;
    const size_t last = (size_t)ctx->cfg.decision_delay + ctx->cfg.q_offset;
    if (L <= last) return 0;
    return (L - 1 - last) / ctx->cfg.samples_per_symbol + 1;
}

int modem_gpu_modulate(modem_ctx_t* ctx, const uint8_t* bits, size_t F, size_t nbits, modem_c32_t* tx, modem_c32_t* iq)
{
    if (!ctx || (!bits && F && nbits) || (!tx && !iq)) return fail(ctx, MODEM_ERR_INVALID, "modulate: null argument");
    CK(ctx, cudaSetDevice(ctx->device));
    if (ctx->n_channels && F > ctx->n_channels * ctx->frames_per_channel)
        return fail(ctx, MODEM_ERR_INVALID, "modulate: more frames than channels * frames_per_channel");
    const size_t L = modem_gpu_frame_samples(ctx, nbits);
    Staged sb, st, si;
    int rc = stage_in(ctx, ctx->s_bits, bits, F * nbits, &sb);
    if (!rc) rc = stage_out(ctx, ctx->s_tx, tx, F * L * sizeof(float2), &st);
    if (!rc) rc = stage_out(ctx, ctx->s_iq, iq, F * L * sizeof(float2), &si);
    if (!rc) rc = launch_tx(ctx, (const uint8_t*)sb.dev, F, nbits, (float2*)st.dev, (float2*)si.dev);
    if (!rc) rc = finish_out(ctx, st);
    if (!rc) rc = finish_out(ctx, si);
    if (!rc && (st.host || si.host)) CK(ctx, cudaStreamSynchronize(ctx->stream));
    return rc;
}

int modem_gpu_set_phasor(modem_ctx_t* ctx, const modem_phasor_t* ph)
{
    if (!ctx) return MODEM_ERR_INVALID;
    ctx->graph_epoch++;
    if (!ph || ph->kind == MODEM_PHASOR_TABLE) {
        ctx->phasor_on = false;
        return MODEM_OK;
    }
    if (ph->struct_size != sizeof(modem_phasor_t)) return fail(ctx, MODEM_ERR_INVALID, "modem_phasor_t size mismatch (ABI)");
    if (ph->kind > MODEM_PHASOR_DMPSK) return fail(ctx, MODEM_ERR_INVALID, "unknown phasor kind");
    if (ph->bits_per_symbol != ctx->cfg.bits_per_symbol)
        return fail(ctx, MODEM_ERR_INVALID, "phasor bits_per_symbol differs from the context's");
    if (ph->kind == MODEM_PHASOR_BFSK && ph->bits_per_symbol != 1) return fail(ctx, MODEM_ERR_INVALID, "bfsk carries 1 bit per symbol (bfsk.rs:33)");
    if (ph->kind == MODEM_PHASOR_MSK) {
        if (ph->bits_per_symbol != 2) return fail(ctx, MODEM_ERR_INVALID, "msk carries 2 bits per symbol (msk.rs:27)");
        if (ctx->cfg.samples_per_symbol % 2) return fail(ctx, MODEM_ERR_INVALID, "assertion failed: samples_per_symbol % 2 == 0 (msk.rs:13)");
    } else if (ctx->cfg.q_offset) {
        return fail(ctx, MODEM_ERR_INVALID, "q_offset (EvenOddOffset) is only defined for 2-bit memoryless mappers and msk");
    }
    ctx->phasor = *ph;
    ctx->phasor_on = true;
    return MODEM_OK;
}

int modem_gpu_preamble(modem_ctx_t* ctx, size_t F, size_t n, float amplitude, modem_c32_t* tx)
{
    if (!ctx || (!tx && F && n)) return fail(ctx, MODEM_ERR_INVALID, "preamble: null argument");
    CK(ctx, cudaSetDevice(ctx->device));
    if (F == 0 || n == 0) return MODEM_OK;
    Staged st;
    int rc = stage_out(ctx, ctx->s_tx, tx, F * n * sizeof(float2), &st);
    if (rc) return rc;
    const unsigned blocks = (unsigned)std::min<u64>(((u64)F * n + mg::kThreads - 1) / mg::kThreads, (u64)ctx->sm_count * 16);
    mg::tone_kernel<<<blocks, mg::kThreads, 0, ctx->stream>>>((float2*)st.dev, nullptr, 0, F, n, amplitude, channel_view(ctx), ctx->cfg.sample0);
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    rc = finish_out(ctx, st);
    if (!rc && st.host) CK(ctx, cudaStreamSynchronize(ctx->stream));
    return rc;
}

int modem_gpu_modulate_real(modem_ctx_t* ctx, const uint8_t* bits, size_t F, size_t nbits, size_t preamble,
                            float preamble_amplitude, float* out)
{
    if (!ctx || (!bits && F && nbits) || !out) return fail(ctx, MODEM_ERR_INVALID, "modulate_real: null argument");
    CK(ctx, cudaSetDevice(ctx->device));
    if (ctx->n_channels && F > ctx->n_channels * ctx->frames_per_channel)
        return fail(ctx, MODEM_ERR_INVALID, "modulate_real: more frames than channels * frames_per_channel");
    const size_t L = modem_gpu_frame_samples(ctx, nbits);
    const size_t row = preamble + L;
    if (F == 0 || row == 0) return MODEM_OK;
    Staged sb, so;
    int rc = stage_in(ctx, ctx->s_bits, bits, F * nbits, &sb);
    if (!rc) rc = stage_out(ctx, ctx->s_re, out, F * row * sizeof(float), &so);
    if (rc) return rc;
    if (preamble) {
        const unsigned blocks = (unsigned)std::min<u64>(((u64)F * preamble + mg::kThreads - 1) / mg::kThreads, (u64)ctx->sm_count * 16);
        mg::tone_kernel<<<blocks, mg::kThreads, 0, ctx->stream>>>(nullptr, (float*)so.dev, row, F, preamble, preamble_amplitude,
                                                                 channel_view(ctx), ctx->cfg.sample0);
        ctx->launches++;
        CK(ctx, cudaGetLastError());
    }
    RealOut ro;
    ro.re = (float*)so.dev;
    ro.stride = row;
    ro.offset = preamble;
    ro.sample_skip = preamble;
    rc = launch_tx(ctx, (const uint8_t*)sb.dev, F, nbits, nullptr, nullptr, &ro);
    if (!rc) rc = finish_out(ctx, so);
    if (!rc && so.host) CK(ctx, cudaStreamSynchronize(ctx->stream));
    return rc;
}

int modem_gpu_lock_phase(modem_ctx_t* ctx, const void* samples, uint32_t fmt, size_t F, size_t L, const float* hilbert_taps,
                         size_t n_hilbert, size_t lock_samples, float* phase_offset)
{
    if (!ctx || (!samples && F && L) || !phase_offset) return fail(ctx, MODEM_ERR_INVALID, "lock_phase: null argument");
    if (fmt > MODEM_SAMPLES_I16) return fail(ctx, MODEM_ERR_INVALID, "lock_phase: unknown sample format");
    if (L < lock_samples) return fail(ctx, MODEM_ERR_INVALID, "lock_phase: called `Option::unwrap()` on a `None` value (fewer samples than the lock needs, demodulator.rs:34)");
    CK(ctx, cudaSetDevice(ctx->device));
    if (F == 0) return MODEM_OK;
    Staged sr, sp;
    /* only the first lock_samples of each row are needed, but rows are contiguous: stage the lot */
    int rc = stage_in(ctx, ctx->s_raw, samples, F * L * sample_bytes(fmt), &sr);
    if (!rc) rc = stage_out(ctx, ctx->s_po, phase_offset, F * sizeof(float), &sp);
    if (!rc) rc = launch_lock(ctx, sr.dev, fmt, F, L, hilbert_taps, n_hilbert, lock_samples, (float*)sp.dev);
    if (!rc) rc = finish_out(ctx, sp);
    if (!rc && sp.host) CK(ctx, cudaStreamSynchronize(ctx->stream));
    return rc;
}

int modem_gpu_demodulate_real(modem_ctx_t* ctx, const void* samples, uint32_t fmt, size_t F, size_t L, size_t lock_samples,
                              const float* hilbert_taps, size_t n_hilbert, float* phase_offset, uint8_t* sym, uint8_t* bits,
                              modem_c32_t* soft, modem_c32_t* filt)
{
    if (!ctx || (!samples && F && L)) return fail(ctx, MODEM_ERR_INVALID, "demodulate_real: null argument");
    if (fmt > MODEM_SAMPLES_I16) return fail(ctx, MODEM_ERR_INVALID, "demodulate_real: unknown sample format");
    if (L < lock_samples) return fail(ctx, MODEM_ERR_INVALID, "demodulate_real: called `Option::unwrap()` on a `None` value (fewer samples than the lock needs, demodulator.rs:34)");
    CK(ctx, cudaSetDevice(ctx->device));
    if (ctx->n_channels && F > ctx->n_channels * ctx->frames_per_channel)
        return fail(ctx, MODEM_ERR_INVALID, "demodulate_real: more frames than channels * frames_per_channel");
    if (F == 0) return MODEM_OK;
    const size_t Lr = L - lock_samples;
    const size_t K = modem_gpu_decided_symbols(ctx, Lr);
    const size_t bps = ctx->cfg.bits_per_symbol;
    Staged sr, sp, ss, sb, so, sf;
    int rc = stage_in(ctx, ctx->s_raw, samples, F * L * sample_bytes(fmt), &sr);
    float* d_po = nullptr;
    if (!rc && lock_samples) {
        if (phase_offset) {
            rc = stage_out(ctx, ctx->s_po, phase_offset, F * sizeof(float), &sp);
            d_po = (float*)sp.dev;
        } else {
            rc = ensure(ctx, ctx->s_po, F * sizeof(float));
            d_po = (float*)ctx->s_po.p;
        }
        if (!rc) rc = launch_lock(ctx, sr.dev, fmt, F, L, hilbert_taps, n_hilbert, lock_samples, d_po);
    }
    if (!rc) rc = stage_out(ctx, ctx->s_sym, sym, F * K, &ss);
    if (!rc) rc = stage_out(ctx, ctx->s_bits_out, bits, F * K * bps, &sb);
    if (!rc) rc = stage_out(ctx, ctx->s_soft, soft, F * K * sizeof(float2), &so);
    if (!rc) rc = stage_out(ctx, ctx->s_filt, filt, F * Lr * sizeof(float2), &sf);
    if (!rc && Lr) {
        RawSrc src;
        src.raw = sr.dev;
        src.fmt = fmt;
        src.stride = L;
        src.skip = lock_samples;
        src.d_po = d_po;
        if (fmt == MODEM_SAMPLES_C32) src.fmt = 3; /* analytic rows supplied by the caller: Demodulator::next reads .re */
        rc = launch_rx(ctx, nullptr, F, Lr, (uint8_t*)ss.dev, (uint8_t*)sb.dev, (float2*)so.dev, (float2*)sf.dev, nullptr, 0, nullptr,
                       0.0f, 0, 0, &src);
    }
    if (!rc) rc = finish_out(ctx, sp);
    if (!rc) rc = finish_out(ctx, ss);
    if (!rc) rc = finish_out(ctx, sb);
    if (!rc) rc = finish_out(ctx, so);
    if (!rc) rc = finish_out(ctx, sf);
    if (!rc && (sp.host || ss.host || sb.host || so.host || sf.host)) CK(ctx, cudaStreamSynchronize(ctx->stream));
    return rc;
}

int modem_gpu_awgn(modem_ctx_t* ctx, modem_c32_t* buf, size_t F, size_t L, float sigma, uint64_t seed, uint64_t frame0)
{
    if (!ctx || (!buf && F && L)) return fail(ctx, MODEM_ERR_INVALID, "awgn: null argument");
    CK(ctx, cudaSetDevice(ctx->device));
    if (F == 0 || L == 0) return MODEM_OK;
    const size_t bytes = F * L * sizeof(float2);
    float2* d = (float2*)buf;
    const bool host = !is_device_ptr(buf);
    if (host) {
        int rc = ensure(ctx, ctx->s_rx, bytes);
        if (rc) return rc;
        d = (float2*)ctx->s_rx.p;
        CK(ctx, cudaMemcpyAsync(d, buf, bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    const mg::Noise nz = mg::make_noise(sigma, seed, frame0);
    const u64 quads = (u64)F * ((L + 3) / 4);
    const unsigned blocks = (unsigned)std::min<u64>((quads + mg::kThreads - 1) / mg::kThreads, (u64)ctx->sm_count * 32);
    mg::awgn_kernel<<<blocks, mg::kThreads, 0, ctx->stream>>>(d, F, L, nz);
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    if (host) {
        CK(ctx, cudaMemcpyAsync(buf, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MODEM_OK;
}

int modem_gpu_random_bits(modem_ctx_t* ctx, uint8_t* bits, size_t F, size_t nbits, uint64_t seed, uint64_t frame0)
{
    if (!ctx || (!bits && F && nbits)) return fail(ctx, MODEM_ERR_INVALID, "random_bits: null argument");
    CK(ctx, cudaSetDevice(ctx->device));
    if (F == 0 || nbits == 0) return MODEM_OK;
    Staged sb;
    int rc = stage_out(ctx, ctx->s_bits, bits, F * nbits, &sb);
    if (rc) return rc;
    const u64 total = (u64)F * ((nbits + 127) / 128);
    const unsigned blocks = (unsigned)std::min<u64>((total + mg::kThreads - 1) / mg::kThreads, (u64)ctx->sm_count * 32);
    mg::random_bits_kernel<<<blocks, mg::kThreads, 0, ctx->stream>>>((uint8_t*)sb.dev, F, nbits, seed, frame0);
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    rc = finish_out(ctx, sb);
    if (!rc && sb.host) CK(ctx, cudaStreamSynchronize(ctx->stream));
    return rc;
}

int modem_gpu_demodulate(modem_ctx_t* ctx, const modem_c32_t* rx, size_t F, size_t L, uint8_t* sym, uint8_t* bits,
                         modem_c32_t* soft, modem_c32_t* filt, float sigma, uint64_t seed, uint64_t frame0)
{
    if (!ctx || (!rx && F && L)) return fail(ctx, MODEM_ERR_INVALID, "demodulate: null argument");
    CK(ctx, cudaSetDevice(ctx->device));
    if (ctx->n_channels && F > ctx->n_channels * ctx->frames_per_channel)
        return fail(ctx, MODEM_ERR_INVALID, "demodulate: more frames than channels * frames_per_channel");
    const size_t K = modem_gpu_decided_symbols(ctx, L);
    const size_t bps = ctx->cfg.bits_per_symbol;
    Staged sr, ss, sb, so, sf;
    int rc = stage_in(ctx, ctx->s_rx, rx, F * L * sizeof(float2), &sr);
    if (!rc) rc = stage_out(ctx, ctx->s_sym, sym, F * K, &ss);
    if (!rc) rc = stage_out(ctx, ctx->s_bits_out, bits, F * K * bps, &sb);
    if (!rc) rc = stage_out(ctx, ctx->s_soft, soft, F * K * sizeof(float2), &so);
    if (!rc) rc = stage_out(ctx, ctx->s_filt, filt, F * L * sizeof(float2), &sf);
    if (!rc)
        rc = launch_rx(ctx, (const float2*)sr.dev, F, L, (uint8_t*)ss.dev, (uint8_t*)sb.dev, (float2*)so.dev, (float2*)sf.dev,
                       nullptr, 0, nullptr, sigma, seed, frame0);
    if (!rc) rc = finish_out(ctx, ss);
    if (!rc) rc = finish_out(ctx, sb);
    if (!rc) rc = finish_out(ctx, so);
    if (!rc) rc = finish_out(ctx, sf);
    if (!rc && (ss.host || sb.host || so.host || sf.host)) CK(ctx, cudaStreamSynchronize(ctx->stream));
    return rc;
}

int modem_gpu_demodulate_count(modem_ctx_t* ctx, const modem_c32_t* rx, size_t F, size_t L, uint8_t* sym, uint8_t* bits,
                               const uint8_t* ref_bits, size_t ref_stride, uint64_t* counters, float sigma, uint64_t seed,
                               uint64_t frame0)
{
    if (!ctx || !rx || !ref_bits || !counters) return fail(ctx, MODEM_ERR_INVALID, "demodulate_count: null argument");
    CK(ctx, cudaSetDevice(ctx->device));
    if (!is_device_ptr(rx) || !is_device_ptr(ref_bits) || !is_device_ptr(counters) || (sym && !is_device_ptr(sym)) ||
        (bits && !is_device_ptr(bits)))
        return fail(ctx, MODEM_ERR_INVALID, "demodulate_count: device pointers required");
    if (ctx->n_channels && F > ctx->n_channels * ctx->frames_per_channel)
        return fail(ctx, MODEM_ERR_INVALID, "demodulate_count: more frames than channels * frames_per_channel");
    if (ref_stride < modem_gpu_decided_symbols(ctx, L) * ctx->cfg.bits_per_symbol)
        return fail(ctx, MODEM_ERR_INVALID, "demodulate_count: ref_stride shorter than the decided bits");
    return launch_rx(ctx, (const float2*)rx, F, L, sym, bits, nullptr, nullptr, ref_bits, ref_stride, (u64*)counters, sigma,
                     seed, frame0);
}

} /* extern "C" */

namespace {
/*
 * Host-buffer loopback, pipelined by role: lanes[0] issues every H2D(bits) back to back, lanes[1] runs the kernels chunk
 * after chunk, lanes[2] issues every D2H(sym, bits); one event per chunk and stage hands it to the next stream.  The
 * call's bits and results live in WHOLE-CALL device buffers (64 + 96 MiB at C2: nothing next to 180 GB of HBM), so no
 * stream ever waits for a buffer to be recycled (round 1 used a ring of 4 chunk slots).  The step moves as many bytes in
 * as out and is bound by the two DMA directions running at once: 1.65 ms at C2 against 1.54 ms for the bare copies in the
 * same 16 + 16 transfers (tools/pcie_floor.py).
 *
 * Each chunk runs TX kernel + RX kernel through a chunk-sized sample buffer, NOT the fused loopback kernel, although that
 * one is cheaper (44 us against 72 us per 256-frame chunk, and no sample traffic at all): with it the step takes 2.13 ms.
 * Measured (profiles/r02_e2e_pipeline.txt): the per-chunk event trace shows every 4 MiB copy, in both directions, taking
 * ~140 us beside the fused kernel and ~93 us (its stand-alone time) beside TX + RX; SM clocks, P-state and link speed are
 * identical; chunk size, chunk ramps, the fused kernel's CTA shape, tensor memory on or off and storing the samples after
 * all change nothing; a device-to-device copy stream running BESIDE the fused pipeline brings it to 1.85 ms.  So the
 * copy engines move data faster while the device memory is kept busy with reads and writes (the TX + RX form streams
 * 4 GB through HBM per step, the fused form nothing) -- a property of the memory system's idle behaviour, not of the
 * kernels -- and the form that is cheaper on the SMs loses on the step.  MODEM_GPU_PIPE_FUSED=1 selects it anyway;
 * MODEM_GPU_PIPE_RAMP=1 adds short chunks at both ends of the call (slower with TX + RX, a wash with the fused kernel).
 */
int loopback_pipelined(modem_ctx* ctx, const uint8_t* bits, size_t F, size_t nbits, float sigma, uint64_t seed,
                       uint64_t frame0, uint8_t* sym, uint8_t* bits_out, uint64_t counters[2], size_t Fc)
{
    const size_t L = modem_gpu_frame_samples(ctx, nbits);
    const size_t K = modem_gpu_decided_symbols(ctx, L);
    const size_t bps = ctx->cfg.bits_per_symbol;
    if (!ctx->lanes_ready) {
        for (auto& ln : ctx->lanes) {
            CK(ctx, cudaStreamCreateWithFlags(&ln.s, cudaStreamNonBlocking));
            CK(ctx, cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
        }
        CK(ctx, cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming));
        ctx->lanes_ready = true;
    }
    /* chunk schedule: [fs, fs + n) */
    std::vector<std::pair<size_t, size_t>> chunks;
    {
        const size_t unit = ctx->n_channels ? std::min(Fc, ctx->frames_per_channel) : 1; /* chunk edges stay on what Fc was aligned to */
        std::vector<size_t> head;
        if (ctx->pipe_ramp == 1 && !ctx->n_channels)
            for (size_t n = std::max<size_t>(Fc / 8, 16); n < Fc; n *= 2) head.push_back(n);
        if (ctx->pipe_ramp >= 2 && !ctx->n_channels && Fc / (size_t)ctx->pipe_ramp >= 8) head.push_back(Fc / (size_t)ctx->pipe_ramp); /* one short chunk at each end */
        size_t ramp = 0;
        for (size_t n : head) ramp += n;
        if (2 * ramp + Fc > F) head.clear(), ramp = 0;
        size_t fs = 0;
        for (size_t n : head) chunks.push_back({fs, n}), fs += n;
        const size_t mid_end = F - ramp;
        while (fs < mid_end) {
            size_t n = std::min(Fc, mid_end - fs);
            if (mid_end - fs - n < Fc / 2 && mid_end - fs - n > 0) n = mid_end - fs; /* no runt chunk before the down-ramp */
            chunks.push_back({fs, n});
            fs += n;
        }
        for (size_t i = head.size(); i-- > 0;) chunks.push_back({fs, head[i]}), fs += head[i];
        (void)unit;
    }
    size_t max_n = 0;
    for (auto& c : chunks) max_n = std::max(max_n, c.second);
    while (ctx->pipe_events.size() < 3 * chunks.size()) {
        cudaEvent_t e;
        CK(ctx, cudaEventCreateWithFlags(&e, ctx->pipe_trace ? cudaEventDefault : cudaEventDisableTiming));
        ctx->pipe_events.push_back(e);
    }
    CK(ctx, cudaMemsetAsync(ctx->d_counters, 0, 2 * sizeof(u64), ctx->stream));
    {
        /* NCO tables for the whole call, built once before the streams fork (every chunk shares them) */
        mg::ChannelView dummy = channel_view(ctx);
        int rc0 = attach_carrier_table(ctx, dummy, F, L, false);
        if (rc0) return rc0;
    }
    /* which kernels: the fused loopback kernel (no sample buffer at all) when the shape allows, else TX + RX per chunk */
    mg::RxArgs probe{};
    ctx->frame_base = 0;
    const bool fused = ctx->pipe_fused &&
                       loop_fused_eligible(ctx, reinterpret_cast<const uint8_t*>(uintptr_t(16)), max_n, nbits, nullptr, sigma, probe);
    int rc = ensure(ctx, ctx->s_bits, F * nbits);
    if (!rc && sym) rc = ensure(ctx, ctx->s_sym, F * K);
    if (!rc && bits_out) rc = ensure(ctx, ctx->s_bits_out, F * K * bps);
    if (!rc && !fused) rc = ensure(ctx, ctx->pipe_tx, max_n * L * sizeof(float2));
    if (rc) return rc;
    uint8_t* d_bits = (uint8_t*)ctx->s_bits.p;
    uint8_t* d_sym = sym ? (uint8_t*)ctx->s_sym.p : nullptr;
    uint8_t* d_out = bits_out ? (uint8_t*)ctx->s_bits_out.p : nullptr;
    CK(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
    for (auto& ln : ctx->lanes) CK(ctx, cudaStreamWaitEvent(ln.s, ctx->ev_start, 0));

    cudaStream_t user_stream = ctx->stream;
    cudaStream_t s_in = ctx->lanes[0].s, s_k = ctx->lanes[1].s, s_out = ctx->lanes[2].s;
    /* MODEM_GPU_PIPE_TRACE=1: the pipeline's own events are created with timing and their times printed after the call
     * (no extra events: extra records on three streams changed what they were meant to show) */
    if (ctx->pipe_trace) {
        if (!ctx->pipe_t0) CK(ctx, cudaEventCreate(&ctx->pipe_t0));
        CK(ctx, cudaEventRecord(ctx->pipe_t0, user_stream)); /* time zero */
    }
    const bool copies_out = K && (sym || bits_out);
    cudaError_t e = cudaSuccess;
    /* every copy-in is enqueued first: the engine runs them back to back whatever the kernels do */
    for (size_t c = 0; c < chunks.size() && e == cudaSuccess; ++c) {
        const size_t fs = chunks[c].first, n = chunks[c].second;
        e = cudaMemcpyAsync(d_bits + fs * nbits, bits + fs * nbits, n * nbits, cudaMemcpyHostToDevice, s_in);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->pipe_events[3 * c], s_in);
    }
    for (size_t c = 0; c < chunks.size() && !rc && e == cudaSuccess; ++c) {
        const size_t fs = chunks[c].first, n = chunks[c].second;
        e = cudaStreamWaitEvent(s_k, ctx->pipe_events[3 * c], 0);
        if (e != cudaSuccess) break;
        ctx->stream = s_k; /* launch_* enqueue on ctx->stream */
        ctx->frame_base = fs;
        if (fused) {
            rc = launch_loop_fused(ctx, d_bits + fs * nbits, n, nbits, nullptr, d_sym ? d_sym + fs * K : nullptr,
                                   d_out ? d_out + fs * K * bps : nullptr, ctx->d_counters, sigma);
            rc = rc == 1 ? MODEM_OK : (rc == 0 ? fail(ctx, MODEM_ERR_UNSUPPORTED, "loopback pipeline: internal: chunk not eligible for the fused kernel") : rc);
        } else {
            rc = launch_tx(ctx, d_bits + fs * nbits, n, nbits, (float2*)ctx->pipe_tx.p, nullptr);
            if (!rc)
                rc = launch_rx(ctx, (const float2*)ctx->pipe_tx.p, n, L, d_sym ? d_sym + fs * K : nullptr, d_out ? d_out + fs * K * bps : nullptr,
                               nullptr, nullptr, d_bits + fs * nbits, nbits, ctx->d_counters, sigma, seed, frame0 + fs);
        }
        ctx->stream = user_stream;
        ctx->frame_base = 0;
        if (rc) break;
        e = cudaEventRecord(ctx->pipe_events[3 * c + 1], s_k);
        if (e == cudaSuccess && copies_out) {
            e = cudaStreamWaitEvent(s_out, ctx->pipe_events[3 * c + 1], 0);
            if (e == cudaSuccess && sym) e = cudaMemcpyAsync(sym + fs * K, d_sym + fs * K, n * K, cudaMemcpyDeviceToHost, s_out);
            if (e == cudaSuccess && bits_out)
                e = cudaMemcpyAsync(bits_out + fs * K * bps, d_out + fs * K * bps, n * K * bps, cudaMemcpyDeviceToHost, s_out);
            if (e == cudaSuccess && ctx->pipe_trace) e = cudaEventRecord(ctx->pipe_events[3 * c + 2], s_out);
        }
    }
    if (e != cudaSuccess && !rc) rc = fail(ctx, MODEM_ERR_CUDA, std::string("loopback pipeline: ") + cudaGetErrorString(e));
    ctx->stream = user_stream;
    ctx->frame_base = 0;
    for (auto& ln : ctx->lanes) {
        CK(ctx, cudaEventRecord(ln.done, ln.s));
        CK(ctx, cudaStreamWaitEvent(user_stream, ln.done, 0));
    }
    if (rc) {
        cudaStreamSynchronize(user_stream);
        return rc;
    }
    u64 h[2] = {0, 0};
    CK(ctx, cudaMemcpyAsync(h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost, user_stream));
    CK(ctx, cudaStreamSynchronize(user_stream));
    if (ctx->pipe_trace) {
        fprintf(stderr, "pipe trace (us from call start, %s): chunk frames  copied-in  computed  copied-out\n", fused ? "fused kernel" : "two kernels");
        for (size_t c = 0; c < chunks.size(); ++c) {
            float t[3] = {0, 0, 0};
            for (int j = 0; j < (copies_out ? 3 : 2); ++j) cudaEventElapsedTime(&t[j], ctx->pipe_t0, ctx->pipe_events[3 * c + j]);
            fprintf(stderr, "  %3zu %5zu  %8.1f %8.1f %8.1f\n", c, chunks[c].second, t[0] * 1e3, t[1] * 1e3, t[2] * 1e3);
        }
    }
    if (counters) {
        counters[0] += h[0];
        counters[1] += h[1];
    }
    return MODEM_OK;
}
} // namespace

extern "C" {

int modem_gpu_ber_sweep(modem_ctx_t* ctx, const uint8_t* bits, size_t F, size_t nbits, size_t n_points, const float* sigmas,
                        uint64_t seed, uint64_t frame0, modem_c32_t* tx, uint64_t* counters)
{
    if (!ctx || !bits || !counters || (!sigmas && n_points)) return fail(ctx, MODEM_ERR_INVALID, "ber_sweep: null argument");
    CK(ctx, cudaSetDevice(ctx->device));
    if (!is_device_ptr(bits) || !is_device_ptr(counters) || (tx && !is_device_ptr(tx)))
        return fail(ctx, MODEM_ERR_INVALID, "ber_sweep: device pointers required");
    if (ctx->n_channels && F > ctx->n_channels * ctx->frames_per_channel)
        return fail(ctx, MODEM_ERR_INVALID, "ber_sweep: more frames than channels * frames_per_channel");
    const size_t L = modem_gpu_frame_samples(ctx, nbits);
    if (F == 0 || L == 0 || n_points == 0) return MODEM_OK;
    float2* d_tx = (float2*)tx;
    if (!d_tx) {
        int rc = ensure(ctx, ctx->s_tx, F * L * sizeof(float2));
        if (rc) return rc;
        d_tx = (float2*)ctx->s_tx.p;
    }
    int rc = launch_tx(ctx, bits, F, nbits, d_tx, nullptr);
    for (size_t p = 0; p < n_points && !rc; ++p)
        rc = launch_rx(ctx, d_tx, F, L, nullptr, nullptr, nullptr, nullptr, bits, nbits, (u64*)counters + 2 * p, sigmas[p], seed + p,
                       frame0);
    return rc;
}

int modem_gpu_loopback_device(modem_ctx_t* ctx, const uint8_t* bits, size_t F, size_t nbits, float sigma, uint64_t seed,
                              uint64_t frame0, modem_c32_t* tx, uint8_t* sym, uint8_t* bits_out, uint64_t* counters)
{
    if (!ctx || !bits || !counters) return fail(ctx, MODEM_ERR_INVALID, "loopback_device: null argument");
    CK(ctx, cudaSetDevice(ctx->device));
    if (!is_device_ptr(bits) || !is_device_ptr(counters) || (tx && !is_device_ptr(tx)) || (sym && !is_device_ptr(sym)) ||
        (bits_out && !is_device_ptr(bits_out)))
        return fail(ctx, MODEM_ERR_INVALID, "loopback_device: device pointers required");
    if (ctx->n_channels && F > ctx->n_channels * ctx->frames_per_channel)
        return fail(ctx, MODEM_ERR_INVALID, "loopback_device: more frames than channels * frames_per_channel");
    const size_t L = modem_gpu_frame_samples(ctx, nbits);
    const size_t K = modem_gpu_decided_symbols(ctx, L);
    const size_t bps = ctx->cfg.bits_per_symbol;
    if (F == 0 || L == 0) return MODEM_OK;
    float2* d_tx = (float2*)tx;
    mg::RxArgs probe{};
    if (!d_tx && !loop_fused_eligible(ctx, bits, F, nbits, nullptr, sigma, probe)) { /* the fused kernel needs no sample buffer */
        int rc = ensure(ctx, ctx->s_tx, F * L * sizeof(float2));
        if (rc) return rc;
        d_tx = (float2*)ctx->s_tx.p;
    }
    /* MODEM_GPU_LOOP_CHUNK > 0 cuts the call into chunks of that many frames and overlaps the TX kernel of
     * chunk c+1 with the RX kernel of chunk c (RX then reads its chunk out of L2).  Measured on B200 this is
     * SLOWER than two whole-buffer launches at every chunk size, so the default is one
     * chunk; the chunked form is kept for configurations that are HBM-bound on the RX side. */
    size_t Fc = ctx->loop_chunk ? ctx->loop_chunk : F;
    if (ctx->n_channels) {
        const size_t fc = ctx->frames_per_channel;
        if (Fc >= fc) Fc -= Fc % fc;
        else while (fc % Fc) --Fc;
    }
    if (!ctx->lanes_ready) {
        for (auto& ln : ctx->lanes) {
            CK(ctx, cudaStreamCreateWithFlags(&ln.s, cudaStreamNonBlocking));
            CK(ctx, cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
        }
        CK(ctx, cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming));
        ctx->lanes_ready = true;
    }
    for (auto& e : ctx->ev_pool)
        if (!e) CK(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    cudaStream_t user = ctx->stream, helper = ctx->lanes[0].s;
    /* tables for the whole call are built on the user stream before the fork */
    {
        mg::ChannelView dummy = channel_view(ctx);
        int rc = attach_carrier_table(ctx, dummy, F, L, false);
        if (rc) return rc;
    }
    auto pipeline = [&](cudaStream_t user) -> int { /* `user` = the stream the TX kernels and the join go to */
        ctx->stream = user;
        cudaError_t e = cudaEventRecord(ctx->ev_start, user);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(helper, ctx->ev_start, 0);
        int rc = MODEM_OK;
        size_t c = 0;
        for (size_t fs = 0; fs < F && !rc && e == cudaSuccess; fs += Fc, ++c) {
            const size_t n = std::min(Fc, F - fs);
            ctx->frame_base = fs;
            /* the headline shape runs as ONE kernel (TX samples made, stored and demodulated in place) */
            rc = launch_loop_fused(ctx, bits + fs * nbits, n, nbits, d_tx ? d_tx + fs * L : nullptr, sym ? sym + fs * K : nullptr,
                                   bits_out ? bits_out + fs * K * bps : nullptr, (u64*)counters, sigma);
            if (rc == 1) {
                rc = MODEM_OK;
                continue;
            }
            if (rc) break;
            if (!d_tx) {
                rc = fail(ctx, MODEM_ERR_INVALID, "loopback_device: internal: no sample buffer for the two-kernel path");
                break;
            }
            rc = launch_tx(ctx, bits + fs * nbits, n, nbits, d_tx + fs * L, nullptr); /* on the user stream */
            if (rc) break;
            cudaEvent_t ev = ctx->ev_pool[c % 8];
            e = cudaEventRecord(ev, user);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(helper, ev, 0);
            if (e != cudaSuccess) break;
            ctx->stream = helper; /* the RX of chunk c overlaps the TX of chunk c + 1 */
            rc = launch_rx(ctx, d_tx + fs * L, n, L, sym ? sym + fs * K : nullptr, bits_out ? bits_out + fs * K * bps : nullptr,
                           nullptr, nullptr, bits + fs * nbits, nbits, (u64*)counters, sigma, seed, frame0 + fs);
            ctx->stream = user;
        }
        ctx->frame_base = 0;
        if (e == cudaSuccess) e = cudaEventRecord(ctx->lanes[0].done, helper);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(user, ctx->lanes[0].done, 0); /* join */
        if (e != cudaSuccess && !rc) rc = fail(ctx, MODEM_ERR_CUDA, std::string("loopback pipeline: ") + cudaGetErrorString(e));
        return rc;
    };
    struct RestoreStream { /* launch_* enqueue on ctx->stream; always hand the caller's stream back */
        modem_ctx* c;
        cudaStream_t s;
        ~RestoreStream() { c->stream = s; c->frame_base = 0; }
    } restore{ctx, user};

    /* Replay a captured graph when the same call repeats (one launch instead of ~2 per chunk). */
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CK(ctx, cudaStreamIsCapturing(user, &cap));
    auto& g = ctx->loop_graph;
    const bool same = g.exec && g.bits == bits && g.tx == d_tx && g.sym == sym && g.out == bits_out && g.cnt == counters &&
                      g.F == F && g.nbits == nbits && g.Fc == Fc && g.sigma == sigma && g.seed == seed && g.frame0 == frame0 &&
                      g.chan_version == ctx->chan_version && g.epoch == ctx->graph_epoch;
    if (ctx->use_graph && cap == cudaStreamCaptureStatusNone && F > Fc) {
        if (!same) {
            if (g.exec) {
                cudaGraphExecDestroy(g.exec);
                g.exec = nullptr;
            }
            /* first run eagerly (allocations, function attributes), then capture the same sequence */
            int rc = pipeline(user);
            if (rc) return rc;
            /* capture on an internal stream (the caller's may be the legacy default stream, which cannot
             * capture); the instantiated graph is then launched into the caller's stream */
            cudaStream_t cs = ctx->lanes[1].s;
            cudaGraph_t graph = nullptr;
            CK(ctx, cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
            const uint64_t launches_before = ctx->launches;
            rc = pipeline(cs);
            cudaError_t e = cudaStreamEndCapture(cs, &graph);
            ctx->launches = launches_before; /* captured, not launched */
            if (rc || e != cudaSuccess || !graph) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                return rc ? rc : MODEM_OK; /* the eager run above already did the work */
            }
            e = cudaGraphInstantiate(&g.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) {
                g.exec = nullptr;
                cudaGetLastError();
                return MODEM_OK;
            }
            g.bits = bits; g.tx = d_tx; g.sym = sym; g.out = bits_out; g.cnt = counters;
            g.F = F; g.nbits = nbits; g.Fc = Fc; g.sigma = sigma; g.seed = seed; g.frame0 = frame0;
            g.chan_version = ctx->chan_version;
            g.epoch = ctx->graph_epoch;
            return MODEM_OK;
        }
        CK(ctx, cudaGraphLaunch(g.exec, user));
        ctx->launches += 2 * ((F + Fc - 1) / Fc);
        return MODEM_OK;
    }
    return pipeline(user);
}

int modem_gpu_loopback(modem_ctx_t* ctx, const uint8_t* bits, size_t F, size_t nbits, float sigma, uint64_t seed,
                       uint64_t frame0, modem_c32_t* tx, uint8_t* sym, uint8_t* bits_out, uint64_t counters[2])
{
    if (!ctx || (!bits && F && nbits)) return fail(ctx, MODEM_ERR_INVALID, "loopback: null argument");
    CK(ctx, cudaSetDevice(ctx->device));
    if (ctx->n_channels && F > ctx->n_channels * ctx->frames_per_channel)
        return fail(ctx, MODEM_ERR_INVALID, "loopback: more frames than channels * frames_per_channel");
    const size_t L = modem_gpu_frame_samples(ctx, nbits);
    const size_t K = modem_gpu_decided_symbols(ctx, L);
    const size_t bps = ctx->cfg.bits_per_symbol;
    if (!tx && L && !is_device_ptr(bits) && !is_device_ptr(sym) && !is_device_ptr(bits_out)) {
        /* host buffers and no TX dump requested: chunked pipeline (copy-in, kernels, copy-out streams) */
        size_t Fc = ctx->pipe_chunk ? ctx->pipe_chunk : std::max<size_t>(1, ((size_t)128 << 20) / (L * sizeof(float2)));
        if (ctx->n_channels) { /* chunk boundaries on channel boundaries (or whole divisors of them) */
            const size_t fc = ctx->frames_per_channel;
            if (Fc >= fc) Fc -= Fc % fc;
            else while (fc % Fc) --Fc;
        }
        if (F > Fc) return loopback_pipelined(ctx, bits, F, nbits, sigma, seed, frame0, sym, bits_out, counters, Fc);
    }
    Staged sb, st, ss, so;
    int rc = stage_in(ctx, ctx->s_bits, bits, F * nbits, &sb);
    if (rc) return rc;
    if (tx) {
        rc = stage_out(ctx, ctx->s_tx, tx, F * L * sizeof(float2), &st);
    } else {
        rc = ensure(ctx, ctx->s_tx, F * L * sizeof(float2));
        st.dev = ctx->s_tx.p;
    }
    if (!rc) rc = stage_out(ctx, ctx->s_sym, sym, F * K, &ss);
    if (!rc) rc = stage_out(ctx, ctx->s_bits_out, bits_out, F * K * bps, &so);
    if (rc) return rc;
    CK(ctx, cudaMemsetAsync(ctx->d_counters, 0, 2 * sizeof(u64), ctx->stream));
    rc = launch_tx(ctx, (const uint8_t*)sb.dev, F, nbits, (float2*)st.dev, nullptr);
    if (!rc)
        rc = launch_rx(ctx, (const float2*)st.dev, F, L, (uint8_t*)ss.dev, (uint8_t*)so.dev, nullptr, nullptr,
                       (const uint8_t*)sb.dev, nbits, ctx->d_counters, sigma, seed, frame0);
    if (!rc) rc = finish_out(ctx, st);
    if (!rc) rc = finish_out(ctx, ss);
    if (!rc) rc = finish_out(ctx, so);
    if (rc) return rc;
    u64 h[2] = {0, 0};
    CK(ctx, cudaMemcpyAsync(h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (counters) {
        counters[0] += h[0];
        counters[1] += h[1];
    }
    return MODEM_OK;
}

/*
 * loopback_packed (extension): the loopback on PACKED payloads, 8 bits per byte in each direction instead of the
 * reference's one byte per bit (data.rs:35-40) -- 1/8 of the bytes over PCIe, which is what bounds modem_gpu_loopback with
 * host buffers.  Per chunk of frames: copy-in (host callers), unpack_bits_kernel into the byte-per-bit rows the path
 * consumes, the loopback kernels (the fused kernel WITHOUT a sample buffer when the shape allows, else TX + RX through a
 * chunk-sized one), pack_bits_kernel, copy-out; three streams by role as in loopback_pipelined.  Device callers run the
 * same sequence in place, one chunk, no copies.  The bit decisions are those of modem_gpu_loopback, bit for bit.
 */
int modem_gpu_loopback_packed(modem_ctx_t* ctx, const uint8_t* packed, size_t F, size_t nbits, float sigma, uint64_t seed,
                              uint64_t frame0, uint8_t* packed_out, uint64_t counters[2])
{
    if (!ctx || (!packed && F && nbits)) return fail(ctx, MODEM_ERR_INVALID, "loopback_packed: null argument");
    CK(ctx, cudaSetDevice(ctx->device));
    if (ctx->n_channels && F > ctx->n_channels * ctx->frames_per_channel)
        return fail(ctx, MODEM_ERR_INVALID, "loopback_packed: more frames than channels * frames_per_channel");
    const size_t L = modem_gpu_frame_samples(ctx, nbits);
    const size_t K = modem_gpu_decided_symbols(ctx, L);
    const size_t bps = ctx->cfg.bits_per_symbol;
    const size_t nout = K * bps, PB = (nbits + 7) / 8, OB = (nout + 7) / 8;
    if (F == 0 || L == 0) return MODEM_OK;
    const bool dev_in = is_device_ptr(packed), dev_out = packed_out && is_device_ptr(packed_out);
    const bool copy_in = !dev_in, copy_out = packed_out && !dev_out && OB;
    size_t Fc = F;
    if (copy_in || copy_out) {
        Fc = ctx->packed_chunk ? ctx->packed_chunk : std::max<size_t>(1, ((size_t)512 << 20) / (L * sizeof(float2)));
        if (ctx->n_channels) { /* chunk boundaries on channel boundaries (or whole divisors of them) */
            const size_t fc = ctx->frames_per_channel;
            if (Fc >= fc) Fc -= Fc % fc;
            else while (fc % Fc) --Fc;
        }
        Fc = std::min(Fc, F);
    }
    const size_t n_chunks = (F + Fc - 1) / Fc;
    if (!ctx->lanes_ready) {
        for (auto& ln : ctx->lanes) {
            CK(ctx, cudaStreamCreateWithFlags(&ln.s, cudaStreamNonBlocking));
            CK(ctx, cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
        }
        CK(ctx, cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming));
        ctx->lanes_ready = true;
    }
    while (ctx->pipe_events.size() < 3 * n_chunks) {
        cudaEvent_t e;
        CK(ctx, cudaEventCreateWithFlags(&e, ctx->pipe_trace ? cudaEventDefault : cudaEventDisableTiming));
        ctx->pipe_events.push_back(e);
    }
    /* does a chunk run the fused kernel (no sample buffer) or TX + RX (one chunk of samples)? */
    mg::RxArgs probe{};
    ctx->frame_base = 0;
    const bool fused = loop_fused_eligible(ctx, reinterpret_cast<const uint8_t*>(uintptr_t(16)), Fc, nbits, nullptr, sigma, probe);
    int rc = ensure(ctx, ctx->s_bits, F * nbits);
    if (!rc && packed_out) rc = ensure(ctx, ctx->s_bits_out, F * nout);
    if (!rc && copy_in) rc = ensure(ctx, ctx->s_packed_in, F * PB);
    if (!rc && copy_out) rc = ensure(ctx, ctx->s_packed_out, F * OB);
    if (!rc && !fused) rc = ensure(ctx, ctx->pipe_tx, Fc * L * sizeof(float2));
    if (rc) return rc;
    uint8_t* d_bits = (uint8_t*)ctx->s_bits.p;
    uint8_t* d_out = packed_out ? (uint8_t*)ctx->s_bits_out.p : nullptr;
    const uint8_t* d_pin = copy_in ? (const uint8_t*)ctx->s_packed_in.p : packed;
    uint8_t* d_pout = copy_out ? (uint8_t*)ctx->s_packed_out.p : packed_out;
    CK(ctx, cudaMemsetAsync(ctx->d_counters, 0, 2 * sizeof(u64), ctx->stream));
    {
        /* NCO tables for the whole call, built once before the streams fork (every chunk shares them) */
        mg::ChannelView dummy = channel_view(ctx);
        int rc0 = attach_carrier_table(ctx, dummy, F, L, false);
        if (rc0) return rc0;
    }
    CK(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
    for (auto& ln : ctx->lanes) CK(ctx, cudaStreamWaitEvent(ln.s, ctx->ev_start, 0));
    cudaStream_t user_stream = ctx->stream;
    cudaStream_t s_in = ctx->lanes[0].s, s_k = ctx->lanes[1].s, s_out = ctx->lanes[2].s;
    cudaError_t e = cudaSuccess;
    if (copy_in)
        for (size_t c = 0; c < n_chunks && e == cudaSuccess; ++c) {
            const size_t fs = c * Fc, n = std::min(Fc, F - fs);
            e = cudaMemcpyAsync(const_cast<uint8_t*>(d_pin) + fs * PB, packed + fs * PB, n * PB, cudaMemcpyHostToDevice, s_in);
            if (e == cudaSuccess) e = cudaEventRecord(ctx->pipe_events[3 * c], s_in);
        }
    auto grid_for = [&](u64 total) { return (unsigned)std::min<u64>((total + mg::kThreads - 1) / mg::kThreads, (u64)ctx->sm_count * 32); };
    for (size_t c = 0; c < n_chunks && !rc && e == cudaSuccess; ++c) {
        const size_t fs = c * Fc, n = std::min(Fc, F - fs);
        if (copy_in) e = cudaStreamWaitEvent(s_k, ctx->pipe_events[3 * c], 0);
        if (e != cudaSuccess) break;
        mg::unpack_bits_kernel<<<grid_for((u64)n * PB), mg::kThreads, 0, s_k>>>(d_pin + fs * PB, d_bits + fs * nbits, n, nbits);
        ctx->launches++;
        ctx->stream = s_k; /* launch_* enqueue on ctx->stream */
        ctx->frame_base = fs;
        rc = launch_loop_fused(ctx, d_bits + fs * nbits, n, nbits, nullptr, nullptr, d_out ? d_out + fs * nout : nullptr, ctx->d_counters, sigma);
        if (rc == 1) {
            rc = MODEM_OK;
        } else if (rc == 0) {
            if (fused) { /* the probe said fused: no sample buffer was reserved */
                rc = fail(ctx, MODEM_ERR_UNSUPPORTED, "loopback_packed: internal: chunk not eligible for the fused kernel");
            } else {
                rc = launch_tx(ctx, d_bits + fs * nbits, n, nbits, (float2*)ctx->pipe_tx.p, nullptr);
                if (!rc)
                    rc = launch_rx(ctx, (const float2*)ctx->pipe_tx.p, n, L, nullptr, d_out ? d_out + fs * nout : nullptr, nullptr, nullptr,
                                   d_bits + fs * nbits, nbits, ctx->d_counters, sigma, seed, frame0 + fs);
            }
        }
        ctx->stream = user_stream;
        ctx->frame_base = 0;
        if (rc) break;
        if (d_pout && OB) {
            mg::pack_bits_kernel<<<grid_for((u64)n * OB), mg::kThreads, 0, s_k>>>(d_out + fs * nout, d_pout + fs * OB, n, nout);
            ctx->launches++;
        }
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaEventRecord(ctx->pipe_events[3 * c + 1], s_k);
        if (e == cudaSuccess && copy_out) {
            e = cudaStreamWaitEvent(s_out, ctx->pipe_events[3 * c + 1], 0);
            if (e == cudaSuccess) e = cudaMemcpyAsync(packed_out + fs * OB, d_pout + fs * OB, n * OB, cudaMemcpyDeviceToHost, s_out);
        }
    }
    if (e != cudaSuccess && !rc) rc = fail(ctx, MODEM_ERR_CUDA, std::string("loopback_packed: ") + cudaGetErrorString(e));
    ctx->stream = user_stream;
    ctx->frame_base = 0;
    for (auto& ln : ctx->lanes) {
        CK(ctx, cudaEventRecord(ln.done, ln.s));
        CK(ctx, cudaStreamWaitEvent(user_stream, ln.done, 0));
    }
    if (rc) {
        cudaStreamSynchronize(user_stream);
        return rc;
    }
    u64 h[2] = {0, 0};
    CK(ctx, cudaMemcpyAsync(h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost, user_stream));
    CK(ctx, cudaStreamSynchronize(user_stream));
    if (counters) {
        counters[0] += h[0];
        counters[1] += h[1];
    }
    return MODEM_OK;
}

/* ------------------------------------------------------------------ memory helpers */
int modem_gpu_malloc(modem_ctx_t* ctx, void** dptr, size_t bytes)
{
    if (!ctx || !dptr) return MODEM_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, MODEM_ERR_NOMEM, "cudaMalloc failed");
    }
    return MODEM_OK;
}
int modem_gpu_free(modem_ctx_t* ctx, void* dptr)
{
    if (!ctx) return MODEM_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaFree(dptr));
    return MODEM_OK;
}
int modem_gpu_host_alloc(void** hptr, size_t bytes)
{
    if (!hptr) return MODEM_ERR_INVALID;
    cudaError_t e = cudaMallocHost(hptr, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, MODEM_ERR_NOMEM, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
    }
    return MODEM_OK;
}
int modem_gpu_host_free(void* hptr)
{
    CK(nullptr, cudaFreeHost(hptr));
    return MODEM_OK;
}
int modem_gpu_memcpy_h2d(modem_ctx_t* ctx, void* dst, const void* src, size_t bytes)
{
    if (!ctx) return MODEM_ERR_INVALID;
    CK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return MODEM_OK;
}
int modem_gpu_memcpy_d2h(modem_ctx_t* ctx, void* dst, const void* src, size_t bytes)
{
    if (!ctx) return MODEM_ERR_INVALID;
    CK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return MODEM_OK;
}

/* ------------------------------------------------------------------ NCCL (lazy) */
struct Id128 { /* ncclUniqueId */
    char internal[MODEM_COMM_ID_BYTES];
};
namespace {
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128 /* by value */, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
int nccl_load()
{
    if (g_nccl.h) return MODEM_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h) break;
    }
    if (!g_nccl.h) return fail(nullptr, MODEM_ERR_NCCL, std::string("dlopen libnccl.so.2: ") + dlerror());
    g_nccl.GetUniqueId = (int (*)(void*))dlsym(g_nccl.h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(g_nccl.h, "ncclCommInitRank");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.h, "ncclAllReduce");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(g_nccl.h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
        return fail(nullptr, MODEM_ERR_NCCL, "libnccl is missing required symbols");
    return MODEM_OK;
}
int nccl_fail(modem_ctx* ctx, const char* what, int r)
{
    return fail(ctx, MODEM_ERR_NCCL, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
}
} // namespace

int modem_gpu_comm_unique_id(uint8_t id[MODEM_COMM_ID_BYTES])
{
    if (!id) return MODEM_ERR_INVALID;
    int rc = nccl_load();
    if (rc) return rc;
    Id128 u;
    memset(&u, 0, sizeof u);
    int r = g_nccl.GetUniqueId(&u);
    if (r) return nccl_fail(nullptr, "ncclGetUniqueId", r);
    memcpy(id, &u, sizeof u);
    return MODEM_OK;
}

int modem_gpu_comm_create(modem_comm_t** out, modem_ctx_t* ctx, int n_ranks, int rank, const uint8_t id[MODEM_COMM_ID_BYTES])
{
    if (!out || !ctx || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(ctx, MODEM_ERR_INVALID, "comm_create: bad arguments");
    int rc = nccl_load();
    if (rc) return rc;
    CK(ctx, cudaSetDevice(ctx->device));
    modem_comm* c = new modem_comm();
    c->ctx = ctx;
    Id128 u;
    memcpy(&u, id, sizeof u);
    int r = g_nccl.CommInitRank(&c->nccl_comm, n_ranks, u, rank);
    if (r) {
        delete c;
        return nccl_fail(ctx, "ncclCommInitRank", r);
    }
    *out = c;
    return MODEM_OK;
}

int modem_gpu_allreduce_counters(modem_comm_t* comm, uint64_t* counters, size_t n)
{
    if (!comm || !counters || n == 0) return MODEM_ERR_INVALID;
    modem_ctx* ctx = comm->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    const bool dev = is_device_ptr(counters);
    u64* d = (u64*)counters;
    if (!dev) {
        if (comm->cap < n) {
            if (comm->d_buf) CK(ctx, cudaFree(comm->d_buf));
            CK(ctx, cudaMalloc((void**)&comm->d_buf, n * sizeof(u64)));
            comm->cap = n;
        }
        d = comm->d_buf;
        CK(ctx, cudaMemcpyAsync(d, counters, n * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    }
    /* the single collective of the path: ncclUint64 = 5, ncclSum = 0 */
    int r = g_nccl.AllReduce(d, d, n, 5, 0, comm->nccl_comm, ctx->stream);
    if (r) return nccl_fail(ctx, "ncclAllReduce", r);
    if (!dev) {
        CK(ctx, cudaMemcpyAsync(counters, d, n * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MODEM_OK;
}

void modem_gpu_comm_destroy(modem_comm_t* comm)
{
    if (!comm) return;
    if (comm->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm->nccl_comm);
    if (comm->d_buf) cudaFree(comm->d_buf);
    delete comm;
}

} /* extern "C" */
