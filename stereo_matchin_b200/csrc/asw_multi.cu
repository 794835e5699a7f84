// asw_multi.cu -- one frame on several GPUs of one process, through the C ABI only.
//
// The reference enumerates every OpenCL device and runs the whole job on each of them in turn
// (stereo_matching/main.cpp:119-130,158-172): it has no multi-device split.  Here the devices share ONE frame:
// row bands, one host thread and one asw_ctx per GPU, and between two aggregation iterations every band pulls the
// `radius` boundary rows of its neighbours straight out of their cost volumes over NVLink (cudaMemcpyPeerAsync).
// Nothing is recomputed (the r * R halo of asw_disparity_band_device is replaced by R exchanged rows), so the
// result is bit-identical to the one-GPU frame.  No collective library is involved: the only exchange of the path is
// these neighbour copies, and every band writes its rows of the result directly into the caller's host buffers.
// The copies run on a communication stream per band, ordered by CUDA events across the devices, under the interior
// rows of the iteration (asw_disparity_band_exchange_async_device); the band threads only meet at host barriers that
// make sure an event has been RECORDED before a neighbour enqueues a wait on it -- no thread waits for GPU work.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/asw_b200.h"

namespace {

// barrier for the band threads; wait() returns false once any thread has called fail() (nobody blocks after an error)
struct Barrier {
    std::mutex m;
    std::condition_variable cv;
    int n = 0, count = 0, gen = 0;
    bool broken = false;
    bool wait() {
        std::unique_lock<std::mutex> l(m);
        if (broken) return false;
        const int g = gen;
        if (++count == n) { count = 0; gen++; cv.notify_all(); return true; }
        cv.wait(l, [&] { return gen != g || broken; });
        return !broken;
    }
    void fail() {
        std::lock_guard<std::mutex> l(m);
        broken = true;
        cv.notify_all();
    }
};

struct BandState {
    asw_ctx* ctx = nullptr;
    int device = 0;
    void* buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // left, right, out_rgba, out_d, out_conf (device, grown on demand)
    size_t cap[5] = {0, 0, 0, 0, 0};
    void *top_send = nullptr, *bottom_send = nullptr;      // published at every exchange
    cudaStream_t comm = nullptr;                           // the band's communication stream
    cudaEvent_t ready = nullptr, done = nullptr;           // boundary rows of this iteration final / this band's pulls complete
    asw_timing tm;
    int status = ASW_OK;
};

}  // namespace

struct asw_multi {
    int n = 0;
    std::vector<BandState> band;
    Barrier bar;
    std::string err;
};

namespace {

struct CbArg {
    asw_multi* m;
    int i;
};

// asw_halo_begin_fn of band i (contract: asw_disparity_band_exchange_async_device).  `boundary_stream` has the band's
// border rows of this iteration enqueued.
int begin_cb(void* user, int, void* top_send, void* bottom_send, void* top_recv, void* bottom_recv, size_t bytes, void* boundary_stream) {
    CbArg* a = (CbArg*)user;
    asw_multi* m = a->m;
    const int i = a->i;
    BandState& b = m->band[i];
    b.top_send = top_send;
    b.bottom_send = bottom_send;
    if (cudaEventRecord(b.ready, (cudaStream_t)boundary_stream) != cudaSuccess) { m->bar.fail(); return 1; }
    if (!m->bar.wait()) return 1;                          // every band has recorded `ready` and published its rows
    cudaError_t e = cudaStreamWaitEvent(b.comm, b.ready, 0);               // our halo rows are no longer read by our vertical pass
    if (e == cudaSuccess && top_recv && i > 0) {
        e = cudaStreamWaitEvent(b.comm, m->band[i - 1].ready, 0);
        if (e == cudaSuccess) e = cudaMemcpyPeerAsync(top_recv, b.device, m->band[i - 1].bottom_send, m->band[i - 1].device, bytes, b.comm);
    }
    if (e == cudaSuccess && bottom_recv && i + 1 < m->n) {
        e = cudaStreamWaitEvent(b.comm, m->band[i + 1].ready, 0);
        if (e == cudaSuccess) e = cudaMemcpyPeerAsync(bottom_recv, b.device, m->band[i + 1].top_send, m->band[i + 1].device, bytes, b.comm);
    }
    if (e == cudaSuccess) e = cudaEventRecord(b.done, b.comm);
    if (e != cudaSuccess) { m->bar.fail(); return 1; }
    return m->bar.wait() ? 0 : 1;                          // every band has recorded `done`
}

// asw_halo_end_fn: the main stream waits for our pulls (halo rows complete) and for the neighbours' pulls (our send rows are
// free: the next horizontal pass rewrites them)
int end_cb(void* user, int, void* main_stream) {
    CbArg* a = (CbArg*)user;
    asw_multi* m = a->m;
    const int i = a->i;
    cudaStream_t s = (cudaStream_t)main_stream;
    cudaError_t e = cudaStreamWaitEvent(s, m->band[i].done, 0);
    if (e == cudaSuccess && i > 0) e = cudaStreamWaitEvent(s, m->band[i - 1].done, 0);
    if (e == cudaSuccess && i + 1 < m->n) e = cudaStreamWaitEvent(s, m->band[i + 1].done, 0);
    if (e != cudaSuccess) { m->bar.fail(); return 1; }
    return 0;
}

int ensure_dev(BandState& b, int k, size_t bytes) {
    if (b.cap[k] >= bytes) return ASW_OK;
    if (b.buf[k]) asw_dev_free(b.ctx, b.buf[k]);
    b.buf[k] = nullptr;
    b.cap[k] = 0;
    int st = asw_dev_alloc(b.ctx, &b.buf[k], bytes);
    if (st == ASW_OK) b.cap[k] = bytes;
    return st;
}

}  // namespace

extern "C" {

int asw_multi_create(asw_multi** out, const int* devices, int n) {
    if (!out || !devices || n <= 0) return ASW_ERR_INVALID;
    *out = nullptr;
    asw_multi* m = new (std::nothrow) asw_multi();
    if (!m) return ASW_ERR_NOMEM;
    m->n = n;
    m->band.resize(n);
    m->bar.n = n;
    for (int i = 0; i < n; i++) {
        m->band[i].device = devices[i];
        int st = asw_create(&m->band[i].ctx, devices[i]);
        if (st != ASW_OK) { asw_multi_destroy(m); return st; }
        BandState& b = m->band[i];
        if (cudaSetDevice(devices[i]) != cudaSuccess || cudaStreamCreateWithFlags(&b.comm, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&b.ready, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming) != cudaSuccess) {
            asw_multi_destroy(m);
            return ASW_ERR_CUDA;
        }
    }
    // neighbours read each other's volumes: peer access makes the copies direct NVLink transfers (without it the
    // runtime stages them through the host, which is still correct)
    for (int i = 0; i < n; i++)
        for (int j = i - 1; j <= i + 1; j += 2) {
            if (j < 0 || j >= n || devices[i] == devices[j]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
            if (can) {
                cudaSetDevice(devices[i]);
                cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
                if (e != cudaSuccess) cudaGetLastError();      // already enabled: fine
            }
        }
    *out = m;
    return ASW_OK;
}

int asw_multi_destroy(asw_multi* m) {
    if (!m) return ASW_ERR_INVALID;
    for (BandState& b : m->band) {
        if (!b.ctx) continue;
        for (void* p : b.buf)
            if (p) asw_dev_free(b.ctx, p);
        cudaSetDevice(b.device);
        if (b.comm) { cudaStreamSynchronize(b.comm); cudaStreamDestroy(b.comm); }
        if (b.ready) cudaEventDestroy(b.ready);
        if (b.done) cudaEventDestroy(b.done);
        asw_destroy(b.ctx);
    }
    delete m;
    return ASW_OK;
}

int asw_multi_count(asw_multi* m) { return m ? m->n : 0; }
const char* asw_multi_last_error(asw_multi* m) { return m ? m->err.c_str() : "null handle"; }

int asw_multi_disparity(asw_multi* m, const uint8_t* left, const uint8_t* right, int W, int H, const asw_params* prm,
                        uint8_t* disp_rgba, uint8_t* disp_d, float* conf, asw_multi_timing* tm) {
    if (!m || !left || !right || !prm || W <= 0 || H <= 0) return ASW_ERR_INVALID;
    const int n = m->n, R = prm->radius;
    if (n > 1 && H / n < R) { m->err = "bands would have fewer than `radius` rows: use fewer devices"; return ASW_ERR_INVALID; }
    {
        std::lock_guard<std::mutex> l(m->bar.m);
        m->bar.broken = false;
        m->bar.count = 0;
    }
    using clk = std::chrono::steady_clock;
    clk::time_point t_begin = clk::now(), t_up, t_done;
    std::vector<std::thread> th;
    std::vector<CbArg> args(n);
    for (int i = 0; i < n; i++) {
        args[i] = CbArg{m, i};
        th.emplace_back([&, i] {
            BandState& b = m->band[i];
            b.status = ASW_OK;
            const int y0 = (int)((long long)H * i / n), y1 = (int)((long long)H * (i + 1) / n);
            const int ya = y0 - R < 0 ? 0 : y0 - R, yb = y1 + R > H ? H : y1 + R;      // rows of the images this band looks at
            const size_t npx = (size_t)W * H, rows = (size_t)(y1 - y0);
            auto bail = [&](int st) { b.status = st; m->bar.fail(); };
            cudaSetDevice(b.device);
            int st;
            const size_t tallest = ((size_t)H / n + 1) * W * 4;
            if ((st = ensure_dev(b, 0, npx * 4)) || (st = ensure_dev(b, 1, npx * 4)) || (st = ensure_dev(b, 2, tallest)) ||
                (st = ensure_dev(b, 3, tallest)) || (st = ensure_dev(b, 4, tallest)))
                return bail(st);
            void *img_l = b.buf[0], *img_r = b.buf[1], *out_rgba = b.buf[2], *out_d = b.buf[3], *out_conf = b.buf[4];
            // only the rows this band reads are uploaded (the band entry point indexes the full frame)
            const size_t off = (size_t)ya * W * 4, len = (size_t)(yb - ya) * W * 4;
            if ((st = asw_memcpy_h2d(b.ctx, (char*)img_l + off, left + off, len)) ||
                (st = asw_memcpy_h2d(b.ctx, (char*)img_r + off, right + off, len)))
                return bail(st);
            if (!m->bar.wait()) return;
            if (i == 0) t_up = clk::now();
            if (n == 1)
                st = asw_disparity_band_device(b.ctx, (const uint8_t*)img_l, (const uint8_t*)img_r, W, H, y0, y1, prm,
                                               disp_rgba ? (uint8_t*)out_rgba : nullptr, disp_d ? (uint8_t*)out_d : nullptr,
                                               conf ? (float*)out_conf : nullptr, &b.tm);
            else
                st = asw_disparity_band_exchange_async_device(b.ctx, (const uint8_t*)img_l, (const uint8_t*)img_r, W, H, y0, y1, prm,
                                                              disp_rgba ? (uint8_t*)out_rgba : nullptr, disp_d ? (uint8_t*)out_d : nullptr,
                                                              conf ? (float*)out_conf : nullptr, begin_cb, end_cb, &args[i], &b.tm);
            if (st) return bail(st);
            if (!m->bar.wait()) return;                        // slowest band: the frame is done on the devices
            if (i == 0) t_done = clk::now();
            if (disp_rgba && (st = asw_memcpy_d2h(b.ctx, disp_rgba + (size_t)y0 * W * 4, out_rgba, rows * W * 4))) return bail(st);
            if (disp_d && (st = asw_memcpy_d2h(b.ctx, disp_d + (size_t)y0 * W, out_d, rows * W))) return bail(st);
            if (conf && (st = asw_memcpy_d2h(b.ctx, conf + (size_t)y0 * W, out_conf, rows * W * 4))) return bail(st);
        });
    }
    for (auto& t : th) t.join();
    clk::time_point t_end = clk::now();
    for (int i = 0; i < n; i++)
        if (m->band[i].status != ASW_OK) {
            char buf[600];
            snprintf(buf, sizeof buf, "band %d (device %d): %s: %s", i, m->band[i].device, asw_strerror(m->band[i].status),
                     asw_last_error(m->band[i].ctx));
            m->err = buf;
            return m->band[i].status;
        }
    if (m->bar.broken) { m->err = "a band failed"; return ASW_ERR_CUDA; }
    if (tm) {
        memset(tm, 0, sizeof *tm);
        auto ms = [](clk::time_point a, clk::time_point b) { return (float)std::chrono::duration<double, std::milli>(b - a).count(); };
        tm->devices = n;
        tm->upload_ms = ms(t_begin, t_up);
        tm->compute_ms = ms(t_up, t_done);
        tm->download_ms = ms(t_done, t_end);
        tm->total_ms = ms(t_begin, t_end);
        for (int i = 0; i < n; i++)
            if (m->band[i].tm.total_ms > tm->slowest_band_device_ms) tm->slowest_band_device_ms = m->band[i].tm.total_ms;
    }
    return ASW_OK;
}

}  // extern "C"
