// api.cu — the C-ABI (include/b3gs.h) and the forward/backward orchestration.
//
// Reference behaviour: rasterizer_impl.cu:197-339 (Rasterizer::forward),
// :343-447 (::backward), :141-153 (::markVisible), rasterizer_impl.h:22-73 (state
// blobs and the obtain/required chunk allocator).
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b3gs.h"
#include "common.cuh"
#include "kernels.h"

namespace b3 {

static std::atomic<unsigned long long> g_launches{0};
void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static thread_local std::string g_error;

static int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
    g_error = what;
    if (e != cudaSuccess) {
        g_error += ": ";
        g_error += cudaGetErrorString(e);
    }
    return code;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- blob layouts: pure functions of (P), (R), (W,H) --------------------------
struct GeomLayout {
    size_t records, depths, tiles_touched, sorted_ids, point_offsets, clamped, rects, counter, grads, scratch, total;
    explicit GeomLayout(int P) {
        const size_t p = (size_t)(P > 0 ? P : 0);
        size_t o = 0;
        records = o;       o = align_up(o + p * 48, 256);
        depths = o;        o = align_up(o + p * 4, 256);
        tiles_touched = o; o = align_up(o + p * 4, 256);
        sorted_ids = o;    o = align_up(o + p * 4, 256);   // Gaussian ids in (depth bits, id) order
        point_offsets = o; o = align_up(o + p * 4, 256);   // exclusive scan of tiles_touched in that order
        clamped = o;       o = align_up(o + p, 256);
        rects = o;         o = align_up(o + p * 8, 256);   // tile rectangles {x0|y0<<16, w|h<<16}
        counter = o;       o = align_up(o + 16, 256);      // R accumulated by the preprocess
        grads = o;         o = align_up(o + p * B3_GRAD_STRIDE * 4, 256);
        // Phase-1 sort scratch is dead once the forward returns; the backward's packed
        // gradient accumulator could alias it, kept separate for clarity (P-sized, small).
        scratch = o;       o = align_up(o + binning_phase1_scratch_bytes(P), 256);
        total = o + 256;
    }
};
struct ImageLayout {
    size_t n_contrib, ranges, total;
    ImageLayout(int W, int H) {
        const size_t n = (size_t)W * H;
        const size_t t = (size_t)((W + B3_TILE_X - 1) / B3_TILE_X) * ((H + B3_TILE_Y - 1) / B3_TILE_Y);
        size_t o = 0;
        n_contrib = o; o = align_up(o + n * 4, 256);
        ranges = o;    o = align_up(o + t * 8, 256);
        total = o + 256;
    }
};
// point_list sits at offset 0 and is all the backward reads, so the backward re-derives what
// it needs from R alone; the forward's scratch behind it depends on the algorithm
// (binning.cu): on (P, tile grid) for the direct tile binning, on R for the radix fallback.
struct BinLayout {
    size_t point_list, scratch, scratch_bytes, total;
    explicit BinLayout(int R) : BinLayout(R, 0, 0, 0, false) {}
    BinLayout(int R, int P, int grid_x, int grid_y, bool with_scratch = true) {
        const size_t r = (size_t)(R > 0 ? R : 0);
        size_t o = 0;
        point_list = o; o = align_up(o + r * 4, 256);
        scratch = o;
        scratch_bytes = with_scratch ? binning_phase2_scratch_bytes(P, R, grid_x, grid_y) : 0;
        o = align_up(o + scratch_bytes, 256);
        total = o + 256;
    }
};

static char* align_ptr(void* p) {
    return reinterpret_cast<char*>(align_up(reinterpret_cast<uintptr_t>(p), 256));
}

// One event and one pinned word per (host thread, device): "R has landed in the pinned word".
// Events belong to the device that was current when they were created, so a thread that
// renders on several GPUs needs one per device.
constexpr int kMaxDevices = 64;
static int current_device() {
    int dev = 0;
    return cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < kMaxDevices ? dev : -1;
}
static cudaEvent_t count_event() {
    static thread_local cudaEvent_t ev[kMaxDevices] = {};
    const int dev = current_device();
    if (dev < 0) return nullptr;
    if (!ev[dev] && cudaEventCreateWithFlags(&ev[dev], cudaEventDisableTiming) != cudaSuccess) ev[dev] = nullptr;
    return ev[dev];
}
static int* pinned_word() {
    static thread_local int* w[kMaxDevices] = {};
    const int dev = current_device();
    if (dev < 0) return nullptr;
    if (!w[dev] && cudaHostAlloc(reinterpret_cast<void**>(&w[dev]), sizeof(int) * 4, cudaHostAllocDefault) != cudaSuccess)
        w[dev] = nullptr;
    return w[dev];
}

// ---- tickets of the no-sync forward -----------------------------------------------------
// b3gs_forward_nosync does not wait for R: it copies the device counter into a pinned slot,
// records an event and hands the slot out as a ticket; b3gs_count_wait(ticket) blocks on that
// event only.  A process-wide ring: a ticket stays valid for kCountSlots further no-sync
// forwards (its generation is checked).
constexpr int kCountSlots = 256;
struct CountSlot {
    cudaEvent_t ev = nullptr;
    int* word = nullptr;        // pinned
    unsigned generation = 0;
};
struct CountRing {              // one per device, allocated as a whole on first use: nothing is
    bool ready = false;         // allocated later, so a forward inside a CUDA-graph capture is legal
    unsigned next = 0;
    CountSlot slots[kCountSlots];
};
static std::mutex g_slot_mu;
static CountRing g_rings[kMaxDevices];

// returns the ticket (device << 24 | generation << 8 | slot) or -1
static int acquire_count_slot(CountSlot** out) {
    const int dev = current_device();
    if (dev < 0) return -1;
    std::lock_guard<std::mutex> lk(g_slot_mu);
    CountRing& r = g_rings[dev];
    if (!r.ready) {
        int* block = nullptr;
        if (cudaHostAlloc(reinterpret_cast<void**>(&block), sizeof(int) * 4 * kCountSlots, cudaHostAllocDefault) != cudaSuccess)
            return -1;
        for (int i = 0; i < kCountSlots; i++) {
            if (cudaEventCreateWithFlags(&r.slots[i].ev, cudaEventDisableTiming) != cudaSuccess) return -1;
            r.slots[i].word = block + 4 * i;
        }
        r.ready = true;
    }
    const unsigned idx = r.next++ % kCountSlots;
    CountSlot& s = r.slots[idx];
    s.generation = (s.generation + 1) & 0xffffu;
    *out = &s;
    return (int)(((unsigned)dev << 24) | (s.generation << 8) | idx);
}

// ---- optional per-stage device timing (bench.py roofline) ---------------------
// When enabled, every stage is bracketed by a pair of CUDA events recorded on the
// caller's stream; b3gs_profile_read() synchronises on them and sums the elapsed
// device time per stage.  Disabled: zero overhead (one relaxed atomic load).
enum Stage { ST_PREPROCESS = 0, ST_SCAN, ST_BINNING, ST_COMPOSITE_FWD, ST_GRAD_ZERO, ST_COMPOSITE_BWD,
             ST_PREPROCESS_BWD, ST_COUNT };
static const char* kStageNames[ST_COUNT] = {"preprocess", "depth_sort", "binning", "composite_forward", "grad_zero",
                                            "composite_backward", "preprocess_backward"};
struct ProfPair { int stage; cudaEvent_t a, b; };
static std::atomic<int> g_profile{0};
static std::mutex g_prof_mu;
static std::vector<ProfPair> g_prof_pending;
static std::vector<cudaEvent_t> g_prof_pool;

static cudaEvent_t prof_event() {
    cudaEvent_t e;
    if (!g_prof_pool.empty()) { e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEventCreate(&e);
    return e;
}
struct StageTimer {
    bool on; int stage; cudaStream_t st; cudaEvent_t a, b;
    StageTimer(int stage_, cudaStream_t st_) : on(g_profile.load(std::memory_order_relaxed) != 0), stage(stage_), st(st_) {
        if (on) {
            std::lock_guard<std::mutex> lk(g_prof_mu);
            a = prof_event(); b = prof_event();
            cudaEventRecord(a, st);
        }
    }
    ~StageTimer() {
        if (on) {
            cudaEventRecord(b, st);
            std::lock_guard<std::mutex> lk(g_prof_mu);
            g_prof_pending.push_back({stage, a, b});
        }
    }
};

#define B3_CHECK_STAGE(what)                                             \
    do {                                                                 \
        cudaError_t e_ = cudaGetLastError();                             \
        if (e_ == cudaSuccess && debug) e_ = cudaStreamSynchronize(st);  \
        if (e_ != cudaSuccess) return fail(B3GS_ERR_CUDA, what, e_);     \
    } while (0)

}  // namespace b3

using namespace b3;

extern "C" {

// capacity < 0: the exact path (one host wait for R).  capacity >= 0: never waits; the binning
// blob holds `capacity` instances, *ticket receives the count ticket.
static int forward_impl(b3gs_buffer geometry, b3gs_buffer binning, b3gs_buffer image, int P, int D, int M,
                        const float* background, int width, int height, const float* means3D, const float* shs,
                        const float* colors_precomp, const float* opacities, const float* scales,
                        float scale_modifier, const float* rotations, const float* cov3D_precomp,
                        const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                        float tan_fovy, int prefiltered, float* out_color, float* out_depth, float* out_alpha,
                        int* radii, int debug, void* stream, int* num_rendered, int capacity, int* ticket,
                        const float* shs_rest = nullptr, int raw = 0) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool nosync = capacity >= 0;
    if (num_rendered) *num_rendered = 0;
    if (P < 0 || width <= 0 || height <= 0 || D < 0 || D > 3)
        return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_forward: bad sizes (P>=0, width,height>0, 0<=D<=3)");
    if (!geometry.resize || !binning.resize || !image.resize)
        return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_forward: null resize callback");
    if (!out_color || !out_depth || !out_alpha || !background)
        return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_forward: null output/background pointer");
    const int grid_x = (width + B3_TILE_X - 1) / B3_TILE_X, grid_y = (height + B3_TILE_Y - 1) / B3_TILE_Y;

    if (P == 0) {
        // rasterize_points.cu:83: the reference skips everything and returns the
        // zero-initialised outputs (NOT the background) and untouched empty blobs.
        const size_t n = (size_t)width * height * sizeof(float);
        cudaError_t e = cudaMemsetAsync(out_color, 0, 3 * n, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(out_depth, 0, n, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(out_alpha, 0, n, st);
        if (e != cudaSuccess) return fail(B3GS_ERR_CUDA, "zero outputs", e);
        return B3GS_OK;
    }

    // image state
    ImageLayout il(width, height);
    char* img = static_cast<char*>(image.resize(image.user, il.total));
    if (!img) return fail(B3GS_ERR_ALLOC, "b3gs_forward: image buffer allocation failed");
    img = align_ptr(img);
    uint32_t* n_contrib = reinterpret_cast<uint32_t*>(img + il.n_contrib);
    uint2* ranges = reinterpret_cast<uint2*>(img + il.ranges);

    int R = 0;
    GeomLayout gl(P);
    char* geo = nullptr;
    if (P > 0) {
        if (!means3D || !opacities || !viewmatrix || !projmatrix || !cam_pos || !radii)
            return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_forward: null required input pointer");
        if ((shs == nullptr) == (colors_precomp == nullptr))
            return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_forward: provide exactly one of shs / colors_precomp");
        if (((scales == nullptr) || (rotations == nullptr)) == (cov3D_precomp == nullptr))
            return fail(B3GS_ERR_INVALID_ARGUMENT,
                        "b3gs_forward: provide exactly one of (scales, rotations) / cov3D_precomp");
        if (shs && M < (D + 1) * (D + 1))
            return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_forward: M < (D+1)^2 SH coefficients");
        geo = static_cast<char*>(geometry.resize(geometry.user, gl.total));
        if (!geo) return fail(B3GS_ERR_ALLOC, "b3gs_forward: geometry buffer allocation failed");
        geo = align_ptr(geo);

        PreprocessArgs pa;
        pa.P = P; pa.D = D; pa.M = M;
        pa.raw = raw; pa.shs_rest = shs_rest;
        pa.means3D = means3D; pa.scales = scales; pa.scale_modifier = scale_modifier; pa.rotations = rotations;
        pa.opacities = opacities; pa.shs = shs; pa.cov3D_precomp = cov3D_precomp; pa.colors_precomp = colors_precomp;
        pa.viewmatrix = viewmatrix; pa.projmatrix = projmatrix; pa.cam_pos = cam_pos;
        pa.W = width; pa.H = height; pa.tan_fovx = tan_fovx; pa.tan_fovy = tan_fovy;
        pa.focal_y = height / (2.0f * tan_fovy);  // rasterizer_impl.cu:223-224
        pa.focal_x = width / (2.0f * tan_fovx);
        pa.grid_x = grid_x; pa.grid_y = grid_y; pa.prefiltered = prefiltered;
        pa.radii = radii;
        pa.records = reinterpret_cast<float4*>(geo + gl.records);
        pa.depths = reinterpret_cast<float*>(geo + gl.depths);
        pa.tiles_touched = reinterpret_cast<uint32_t*>(geo + gl.tiles_touched);
        pa.clamped = reinterpret_cast<uint8_t*>(geo + gl.clamped);
        pa.rects = reinterpret_cast<uint2*>(geo + gl.rects);
        pa.num_rendered = reinterpret_cast<uint32_t*>(geo + gl.counter);
        // counter words: [0] R = 0, [1] OR of depth keys = 0, [2] AND of depth keys = ~0
        cudaError_t e = cudaMemsetAsync(pa.num_rendered, 0, 2 * sizeof(uint32_t), st);
        if (e == cudaSuccess) e = cudaMemsetAsync(pa.num_rendered + 2, 0xff, sizeof(uint32_t), st);
        if (e != cudaSuccess) return fail(B3GS_ERR_CUDA, "zero instance counter", e);
        {
            StageTimer t_(ST_PREPROCESS, st);
            launch_preprocess(pa, st);
        }
        B3_CHECK_STAGE("preprocess");

        // R is needed on the host to size the binning blob (and is part of the operator
        // surface).  Start its copy now, then enqueue the P-sized half of the binning, and
        // only then wait: the GPU sorts by depth while the host sleeps on the event and
        // allocates.  (rasterizer_impl.cu:282 blocks the whole device with cudaMemcpy.)
        int* hw = nullptr;
        cudaEvent_t ev = nullptr;
        if (nosync) {
            CountSlot* slot = nullptr;
            const int tk = acquire_count_slot(&slot);
            if (tk < 0) return fail(B3GS_ERR_ALLOC, "b3gs_forward_nosync: count slot allocation failed");
            hw = slot->word; ev = slot->ev;
            *hw = -1;
            if (ticket) *ticket = tk;
        } else {
            hw = pinned_word();
            ev = count_event();
        }
        if (!hw || !ev) return fail(B3GS_ERR_ALLOC, "b3gs_forward: pinned word / event allocation failed");
        e = cudaMemcpyAsync(hw, pa.num_rendered, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaEventRecord(ev, st);
        if (e != cudaSuccess) return fail(B3GS_ERR_CUDA, "read num_rendered", e);

        {
            BinningPhase1Args b1;
            b1.P = P; b1.depths = pa.depths; b1.tiles_touched = pa.tiles_touched;
            b1.key_bits = pa.num_rendered + 1;
            b1.sorted_ids = reinterpret_cast<uint32_t*>(geo + gl.sorted_ids);
            b1.sorted_offsets = reinterpret_cast<uint32_t*>(geo + gl.point_offsets);
            b1.need_offsets = binning_uses_tile_bins(P, grid_x, grid_y) ? 0 : 1;
            b1.scratch = geo + gl.scratch;
            StageTimer t_(ST_SCAN, st);
            e = run_binning_phase1(b1, st);
        }
        if (e != cudaSuccess) return fail(B3GS_ERR_CUDA, "binning phase 1", e);
        B3_CHECK_STAGE("binning phase 1");

        if (nosync) {
            R = capacity;   // the blob is sized by the caller's estimate; the device clamps to it
        } else {
            e = cudaEventSynchronize(ev);
            if (e != cudaSuccess) return fail(B3GS_ERR_CUDA, "wait for num_rendered", e);
            R = *hw;
        }
    }

    BinLayout bl(R, P, grid_x, grid_y);
    char* bin = static_cast<char*>(binning.resize(binning.user, bl.total));
    if (!bin && bl.total > 0) return fail(B3GS_ERR_ALLOC, "b3gs_forward: binning buffer allocation failed");
    bin = align_ptr(bin);
    uint32_t* point_list = reinterpret_cast<uint32_t*>(bin + bl.point_list);

    {
        BinningPhase2Args ba;
        ba.P = P; ba.R = R; ba.grid_x = grid_x; ba.grid_y = grid_y;
        ba.count_unknown = nosync ? 1 : 0;
        ba.records = reinterpret_cast<const float4*>(geo + gl.records);
        ba.depths = reinterpret_cast<const float*>(geo + gl.depths);
        ba.radii = radii;
        ba.sorted_ids = reinterpret_cast<const uint32_t*>(geo + gl.sorted_ids);
        ba.sorted_offsets = reinterpret_cast<const uint32_t*>(geo + gl.point_offsets);
        ba.rects = reinterpret_cast<const uint2*>(geo + gl.rects);
        ba.point_list = point_list;
        ba.ranges = ranges;
        ba.scratch = bin + bl.scratch;
        cudaError_t e;
        {
            StageTimer t_(ST_BINNING, st);
            e = run_binning_phase2(ba, st);
        }
        if (e != cudaSuccess) return fail(B3GS_ERR_CUDA, "binning", e);
        B3_CHECK_STAGE("binning");
    }

    {
        CompositeFwdArgs ca;
        ca.W = width; ca.H = height; ca.grid_x = grid_x; ca.grid_y = grid_y;
        ca.ranges = ranges; ca.point_list = point_list;
        ca.records = reinterpret_cast<const float4*>(geo + gl.records);
        ca.background = background;
        ca.out_color = out_color; ca.out_depth = out_depth; ca.out_alpha = out_alpha; ca.n_contrib = n_contrib;
        {
            StageTimer t_(ST_COMPOSITE_FWD, st);
            launch_composite_forward(ca, st);
        }
        B3_CHECK_STAGE("composite_forward");
    }
    if (num_rendered) *num_rendered = nosync ? -1 : R;
    return B3GS_OK;
}

int b3gs_forward(b3gs_buffer geometry, b3gs_buffer binning, b3gs_buffer image, int P, int D, int M,
                 const float* background, int width, int height, const float* means3D, const float* shs,
                 const float* colors_precomp, const float* opacities, const float* scales, float scale_modifier,
                 const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                 const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy, int prefiltered,
                 float* out_color, float* out_depth, float* out_alpha, int* radii, int debug, void* stream,
                 int* num_rendered) {
    return forward_impl(geometry, binning, image, P, D, M, background, width, height, means3D, shs, colors_precomp,
                        opacities, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos,
                        tan_fovx, tan_fovy, prefiltered, out_color, out_depth, out_alpha, radii, debug, stream,
                        num_rendered, -1, nullptr);
}

int b3gs_forward_nosync_supported(int P, int width, int height) {
    return P > 0 && binning_uses_tile_bins(P, (width + B3_TILE_X - 1) / B3_TILE_X, (height + B3_TILE_Y - 1) / B3_TILE_Y) ? 1 : 0;
}

int b3gs_forward_nosync(b3gs_buffer geometry, b3gs_buffer binning, b3gs_buffer image, int P, int D, int M,
                        const float* background, int width, int height, const float* means3D, const float* shs,
                        const float* colors_precomp, const float* opacities, const float* scales,
                        float scale_modifier, const float* rotations, const float* cov3D_precomp,
                        const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                        float tan_fovy, int prefiltered, float* out_color, float* out_depth, float* out_alpha,
                        int* radii, void* stream, int capacity, int* ticket) {
    if (capacity < 1 || !ticket) return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_forward_nosync: capacity >= 1 and a ticket are required");
    if (!b3gs_forward_nosync_supported(P, width, height))
        return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_forward_nosync: not supported for these sizes (use b3gs_forward)");
    return forward_impl(geometry, binning, image, P, D, M, background, width, height, means3D, shs, colors_precomp,
                        opacities, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos,
                        tan_fovx, tan_fovy, prefiltered, out_color, out_depth, out_alpha, radii, 0, stream, nullptr,
                        capacity, ticket);
}

int b3gs_count_wait(int ticket, int* num_rendered) {
    if (ticket < 0 || !num_rendered) return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_count_wait: bad ticket");
    cudaEvent_t ev;
    int* word;
    {
        std::lock_guard<std::mutex> lk(g_slot_mu);
        const unsigned dev = (unsigned)ticket >> 24;
        if (dev >= (unsigned)kMaxDevices || !g_rings[dev].ready)
            return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_count_wait: bad ticket");
        const CountSlot& s = g_rings[dev].slots[ticket & 0xff];
        if (s.generation != (((unsigned)ticket >> 8) & 0xffffu))
            return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_count_wait: the ticket has expired (more than 256 later forwards)");
        ev = s.ev; word = s.word;
    }
    cudaError_t e = cudaEventSynchronize(ev);
    if (e != cudaSuccess) return fail(B3GS_ERR_CUDA, "b3gs_count_wait", e);
    *num_rendered = *word;
    return B3GS_OK;
}

static int backward_impl(ExchangePlan* plan, float exchange_scale, unsigned flags, int raw, const float* shs_rest,
                         const float* opacities_raw, float* dL_dsh_rest, int P, int D, int M, int R, const float* background, int width, int height,
                  const float* means3D, const float* shs, const float* colors_precomp, const float* alphas,
                  const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
                  const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx,
                  float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer, char* image_buffer,
                  const float* dL_dpix, const float* dL_dpix_depth, const float* dL_dalphas, float* dL_dmean2D,
                  float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_ddepth, float* dL_dmean3D,
                  float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot, int debug, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (P < 0 || width <= 0 || height <= 0 || R < 0)
        return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_backward: bad sizes");
    if (P == 0) return B3GS_OK;
    if (!geom_buffer || !image_buffer || (!binning_buffer && R > 0))
        return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_backward: null state buffer");
    if (!dL_dpix || !alphas || !radii || !means3D || !background)  // dL_dpix_depth / dL_dalphas: NULL == zeros
        return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_backward: null required input pointer");
    // dL_dconic, dL_dcolor, dL_ddepth, dL_dcov3D may be NULL: intermediates the caller does not want
    if (!dL_dmean2D || !dL_dopacity || !dL_dmean3D || !dL_dscale || !dL_drot || (M > 0 && shs && !dL_dsh))
        return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_backward: null gradient output pointer");
    if (raw && (!opacities_raw || !scales || !rotations || cov3D_precomp || colors_precomp ||
                (M > 1 && (!shs_rest || !dL_dsh_rest))))
        return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_backward_raw: raw parameters missing, or precomputed inputs given");
    const int grid_x = (width + B3_TILE_X - 1) / B3_TILE_X, grid_y = (height + B3_TILE_Y - 1) / B3_TILE_Y;

    GeomLayout gl(P);
    ImageLayout il(width, height);
    BinLayout bl(R);
    char* geo = align_ptr(geom_buffer);
    char* img = align_ptr(image_buffer);
    char* bin = binning_buffer ? align_ptr(binning_buffer) : nullptr;
    float* grads = reinterpret_cast<float*>(geo + gl.grads);

    cudaError_t e;
    {
        StageTimer t_(ST_GRAD_ZERO, st);
        e = cudaMemsetAsync(grads, 0, (size_t)P * B3_GRAD_STRIDE * sizeof(float), st);
    }
    if (e != cudaSuccess) return fail(B3GS_ERR_CUDA, "zero gradient accumulator", e);

    if (R > 0) {
        CompositeBwdArgs ca;
        ca.W = width; ca.H = height; ca.grid_x = grid_x; ca.grid_y = grid_y;
        ca.P = P; ca.R = R;
        ca.ranges = reinterpret_cast<const uint2*>(img + il.ranges);
        ca.point_list = reinterpret_cast<const uint32_t*>(bin + bl.point_list);
        ca.records = reinterpret_cast<const float4*>(geo + gl.records);
        ca.colors_override = nullptr;  // records.c already holds colors_precomp when given
        ca.background = background;
        ca.alphas = alphas;
        ca.n_contrib = reinterpret_cast<const uint32_t*>(img + il.n_contrib);
        ca.dL_dpix = dL_dpix; ca.dL_dpix_depth = dL_dpix_depth; ca.dL_dalphas = dL_dalphas;
        ca.grads = grads;
        {
            StageTimer t_(ST_COMPOSITE_BWD, st);
            launch_composite_backward(ca, st);
        }
        B3_CHECK_STAGE("composite_backward");
    }

    PreBackwardArgs pa;
    pa.P = P; pa.D = D; pa.M = M;
    pa.raw = raw; pa.shs_rest = shs_rest; pa.opacities = opacities_raw; pa.dL_dsh_rest = dL_dsh_rest;
    pa.means3D = means3D; pa.radii = radii; pa.shs = shs;
    pa.clamped = reinterpret_cast<const uint8_t*>(geo + gl.clamped);
    pa.scales = scales; pa.rotations = rotations; pa.scale_modifier = scale_modifier;
    pa.cov3D_precomp = cov3D_precomp; pa.viewmatrix = viewmatrix; pa.projmatrix = projmatrix; pa.campos = campos;
    pa.focal_y = height / (2.0f * tan_fovy);
    pa.focal_x = width / (2.0f * tan_fovx);
    pa.tan_fovx = tan_fovx; pa.tan_fovy = tan_fovy;
    pa.grads = grads;
    pa.dL_dmean2D = dL_dmean2D; pa.dL_dconic = dL_dconic; pa.dL_dopacity = dL_dopacity; pa.dL_dcolor = dL_dcolor;
    pa.dL_ddepth = dL_ddepth; pa.dL_dmean3D = dL_dmean3D; pa.dL_dcov3D = dL_dcov3D; pa.dL_dsh = dL_dsh;
    pa.dL_dscale = dL_dscale; pa.dL_drot = dL_drot;
    pa.accumulate = (flags & B3GS_BWD_ACCUMULATE) ? 1 : 0;
    if (!plan) {
        StageTimer t_(ST_PREPROCESS_BWD, st);
        launch_preprocess_backward(pa, st);
    } else {
        // Backward + exchange, pipelined: K8+K9 runs Gaussian chunk by chunk on the caller's stream
        // and each finished chunk's five (six) gradient ranges are all-reduced over NVLink on the
        // plan's side stream while the next chunk is being computed.
        StageTimer t_(ST_PREPROCESS_BWD, st);
        float* base = exchange_base(plan);
        struct Seg { float* ptr; int width; } segs[6] = {
            {dL_dmean3D, 3}, {dL_dsh, raw ? 3 : 3 * M}, {raw ? dL_dsh_rest : nullptr, 3 * (M - 1)}, {dL_dopacity, 1},
            {dL_dscale, 3}, {dL_drot, 4}};
        for (auto& sg : segs) {
            if (!sg.ptr || sg.width <= 0) { sg.ptr = nullptr; continue; }
            if (sg.ptr < base || ((sg.ptr - base) & 3))
                return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_backward_exchange: gradient outputs must be 16-byte aligned segments of the exchange bucket");
        }
        static const int forced_chunks = [] { const char* e = getenv("B3GS_DP_CHUNKS"); return e ? atoi(e) : 0; }();
        const int chunks = forced_chunks > 0 && forced_chunks <= 16 ? forced_chunks : (P >= (1 << 19) ? 4 : (P >= (1 << 17) ? 2 : 1));
        const int per = (((P + chunks - 1) / chunks) + 3) & ~3;
        cudaStream_t side = exchange_stream(plan);
        int c = 0;
        for (int first = 0; first < P; first += per, c++) {
            const int count = P - first < per ? P - first : per;
            launch_preprocess_backward(pa, st, first, count);
            cudaError_t e2 = cudaEventRecord(exchange_event(plan, c), st);
            if (e2 == cudaSuccess) e2 = cudaStreamWaitEvent(side, exchange_event(plan, c), 0);
            ExchangeRanges rg;
            rg.n = 0;
            for (const auto& sg : segs) {
                if (!sg.ptr) continue;
                const size_t off = (size_t)(sg.ptr - base);
                const size_t a0 = off + (size_t)first * sg.width;                       // multiple of 4: first is
                const size_t a1 = (off + (size_t)(first + count) * sg.width + 3) & ~(size_t)3;  // segment padding is zero
                rg.start4[rg.n] = a0 / 4; rg.len4[rg.n] = (a1 - a0) / 4; rg.n++;
            }
            if (e2 == cudaSuccess) e2 = exchange_chunk(plan, rg, exchange_scale, side);
            if (e2 != cudaSuccess) return fail(B3GS_ERR_CUDA, "b3gs_backward_exchange: chunk exchange", e2);
        }
        cudaError_t e3 = cudaEventRecord(exchange_event(plan, 16), side);
        if (e3 == cudaSuccess) e3 = cudaStreamWaitEvent(st, exchange_event(plan, 16), 0);
        if (e3 != cudaSuccess) return fail(B3GS_ERR_CUDA, "b3gs_backward_exchange: join", e3);
    }
    B3_CHECK_STAGE("preprocess_backward");
    return B3GS_OK;
}

int b3gs_backward_flags(unsigned flags, int P, int D, int M, int R, const float* background, int width, int height,
                        const float* means3D, const float* shs, const float* colors_precomp, const float* alphas,
                        const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
                        const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx,
                        float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer, char* image_buffer,
                        const float* dL_dpix, const float* dL_dpix_depth, const float* dL_dalphas, float* dL_dmean2D,
                        float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_ddepth, float* dL_dmean3D,
                        float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot, int debug, void* stream) {
    return backward_impl(nullptr, 1.0f, flags, 0, nullptr, nullptr, nullptr, P, D, M, R, background, width, height, means3D, shs,
                         colors_precomp, alphas, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix,
                         campos, tan_fovx, tan_fovy, radii, geom_buffer, binning_buffer, image_buffer, dL_dpix,
                         dL_dpix_depth, dL_dalphas, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_ddepth, dL_dmean3D,
                         dL_dcov3D, dL_dsh, dL_dscale, dL_drot, debug, stream);
}

// ---- backward fused with the data-parallel exchange
int b3gs_backward_exchange(void* exchange, float scale, unsigned flags, int P, int D, int M, int R,
                           const float* background, int width, int height, const float* means3D, const float* shs,
                           const float* colors_precomp, const float* alphas, const float* scales, float scale_modifier,
                           const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                           const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy,
                           const int* radii, char* geom_buffer, char* binning_buffer, char* image_buffer,
                           const float* dL_dpix, const float* dL_dpix_depth, const float* dL_dalphas, float* dL_dmean2D,
                           float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_ddepth, float* dL_dmean3D,
                           float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot, int debug, void* stream) {
    if (!exchange) return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_backward_exchange: null exchange handle");
    return backward_impl(static_cast<ExchangePlan*>(exchange), scale, flags, 0, nullptr, nullptr, nullptr, P, D, M, R,
                         background, width, height, means3D, shs, colors_precomp, alphas, scales, scale_modifier,
                         rotations, cov3D_precomp, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, radii, geom_buffer,
                         binning_buffer, image_buffer, dL_dpix, dL_dpix_depth, dL_dalphas, dL_dmean2D, dL_dconic,
                         dL_dopacity, dL_dcolor, dL_ddepth, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, debug,
                         stream);
}

// ---- the raw-parameter entry: activations fused into K1 and K8+K9
int b3gs_forward_raw(b3gs_buffer geometry, b3gs_buffer binning, b3gs_buffer image, int P, int D, int M,
                     const float* background, int width, int height, const float* xyz, const float* f_dc,
                     const float* f_rest, const float* opacity_raw, const float* scaling_raw, float scale_modifier,
                     const float* rotation_raw, const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                     float tan_fovx, float tan_fovy, float* out_color, float* out_depth, float* out_alpha, int* radii,
                     void* stream, int capacity, int* num_rendered_or_ticket) {
    if (!f_dc || !opacity_raw || !scaling_raw || !rotation_raw || M < 1 || (M > 1 && !f_rest))
        return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_forward_raw: f_dc, (f_rest), opacity, scaling, rotation are required");
    if (capacity >= 0 && !b3gs_forward_nosync_supported(P, width, height))
        return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_forward_raw: capacity >= 0 (no-sync) is not supported for these sizes");
    if (capacity >= 0 && (capacity < 1 || !num_rendered_or_ticket))
        return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_forward_raw: capacity >= 1 and a ticket are required");
    return forward_impl(geometry, binning, image, P, D, M, background, width, height, xyz, f_dc, nullptr, opacity_raw,
                        scaling_raw, scale_modifier, rotation_raw, nullptr, viewmatrix, projmatrix, cam_pos, tan_fovx,
                        tan_fovy, 0, out_color, out_depth, out_alpha, radii, 0, stream,
                        capacity < 0 ? num_rendered_or_ticket : nullptr, capacity,
                        capacity < 0 ? nullptr : num_rendered_or_ticket, f_rest, 1);
}

int b3gs_backward_raw(unsigned flags, int P, int D, int M, int R, const float* background, int width, int height,
                      const float* xyz, const float* f_dc, const float* f_rest, const float* opacity_raw,
                      const float* scaling_raw, float scale_modifier, const float* rotation_raw, const float* alphas,
                      const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx,
                      float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer, char* image_buffer,
                      const float* dL_dpix, const float* dL_dpix_depth, const float* dL_dalphas, float* dL_dmean2D,
                      float* dL_dxyz, float* dL_df_dc, float* dL_df_rest, float* dL_dopacity_raw,
                      float* dL_dscaling_raw, float* dL_drotation_raw, void* stream) {
    return backward_impl(nullptr, 1.0f, flags, 1, f_rest, opacity_raw, dL_df_rest, P, D, M, R, background, width, height, xyz, f_dc,
                         nullptr, alphas, scaling_raw, scale_modifier, rotation_raw, nullptr, viewmatrix, projmatrix,
                         campos, tan_fovx, tan_fovy, radii, geom_buffer, binning_buffer, image_buffer, dL_dpix,
                         dL_dpix_depth, dL_dalphas, dL_dmean2D, nullptr, dL_dopacity_raw, nullptr, nullptr, dL_dxyz,
                         nullptr, dL_df_dc, dL_dscaling_raw, dL_drotation_raw, 0, stream);
}

int b3gs_backward(int P, int D, int M, int R, const float* background, int width, int height, const float* means3D,
                  const float* shs, const float* colors_precomp, const float* alphas, const float* scales,
                  float scale_modifier, const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                  const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy, const int* radii,
                  char* geom_buffer, char* binning_buffer, char* image_buffer, const float* dL_dpix,
                  const float* dL_dpix_depth, const float* dL_dalphas, float* dL_dmean2D, float* dL_dconic,
                  float* dL_dopacity, float* dL_dcolor, float* dL_ddepth, float* dL_dmean3D, float* dL_dcov3D,
                  float* dL_dsh, float* dL_dscale, float* dL_drot, int debug, void* stream) {
    return b3gs_backward_flags(0u, P, D, M, R, background, width, height, means3D, shs, colors_precomp, alphas, scales,
                               scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos, tan_fovx,
                               tan_fovy, radii, geom_buffer, binning_buffer, image_buffer, dL_dpix, dL_dpix_depth,
                               dL_dalphas, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_ddepth, dL_dmean3D,
                               dL_dcov3D, dL_dsh, dL_dscale, dL_drot, debug, stream);
}

int b3gs_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* /*projmatrix*/,
                      unsigned char* present, void* stream) {
    if (P < 0) return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_mark_visible: P < 0");
    if (P == 0) return B3GS_OK;
    if (!means3D || !viewmatrix || !present)
        return fail(B3GS_ERR_INVALID_ARGUMENT, "b3gs_mark_visible: null pointer");
    launch_mark_visible(P, means3D, viewmatrix, present, reinterpret_cast<cudaStream_t>(stream));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(B3GS_ERR_CUDA, "mark_visible", e);
    return B3GS_OK;
}

size_t b3gs_geometry_bytes(int P) { return GeomLayout(P).total; }
size_t b3gs_binning_bytes(int R) { return BinLayout(R).total; }
size_t b3gs_binning_bytes_forward(int P, int R, int width, int height) {
    return BinLayout(R, P, (width + B3_TILE_X - 1) / B3_TILE_X, (height + B3_TILE_Y - 1) / B3_TILE_Y).total;
}
size_t b3gs_image_bytes(int width, int height) { return ImageLayout(width, height).total; }

size_t b3gs_geometry_offset(int P, const char* name) {
    GeomLayout l(P);
    if (!strcmp(name, "records")) return l.records;
    if (!strcmp(name, "depths")) return l.depths;
    if (!strcmp(name, "tiles_touched")) return l.tiles_touched;
    if (!strcmp(name, "point_offsets")) return l.point_offsets;
    if (!strcmp(name, "sorted_ids")) return l.sorted_ids;
    if (!strcmp(name, "clamped")) return l.clamped;
    if (!strcmp(name, "rects")) return l.rects;
    if (!strcmp(name, "grads")) return l.grads;
    return (size_t)-1;
}
size_t b3gs_binning_offset(int R, const char* name) {
    BinLayout l(R);
    if (!strcmp(name, "point_list")) return l.point_list;
    return (size_t)-1;
}
size_t b3gs_image_offset(int width, int height, const char* name) {
    ImageLayout l(width, height);
    if (!strcmp(name, "n_contrib")) return l.n_contrib;
    if (!strcmp(name, "ranges")) return l.ranges;
    return (size_t)-1;
}

void b3gs_profile_enable(int on) { g_profile.store(on ? 1 : 0, std::memory_order_relaxed); }
int b3gs_profile_num_stages(void) { return ST_COUNT; }
const char* b3gs_profile_stage_name(int i) { return (i >= 0 && i < ST_COUNT) ? kStageNames[i] : ""; }
int b3gs_profile_read(double* ms_total, unsigned long long* calls, int n) {
    std::vector<ProfPair> pend;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        pend.swap(g_prof_pending);
    }
    for (int i = 0; i < n; i++) { ms_total[i] = 0.0; calls[i] = 0; }
    for (auto& p : pend) {
        float ms = 0.f;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess &&
            p.stage < n) {
            ms_total[p.stage] += ms;
            calls[p.stage] += 1;
        }
    }
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        for (auto& p : pend) { g_prof_pool.push_back(p.a); g_prof_pool.push_back(p.b); }
    }
    return (int)pend.size();
}

void b3gs_set_backward_pixels(int n) { b3::set_backward_pixels(n); }

const char* b3gs_last_error(void) { return g_error.c_str(); }
const char* b3gs_version(void) { return "b3gs 0.1 (sm_100a)"; }
unsigned long long b3gs_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
