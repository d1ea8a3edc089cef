// Result transfers: device -> pinned host copies on a dedicated copy stream, ordered after the stream
// that produced the data.  One C call per copy replaces ~40 us of Python-side event / stream
// bookkeeping on the end-to-end path (the reference returns host arrays: Experiment.py:405, :526,
// main.py:99).  A "lane" owns its two events; the copy stream is shared per device.
#include <mutex>
#include <vector>

#include "common.cuh"

using namespace paresis;

namespace {
struct Lane {
    cudaEvent_t ready = nullptr, done = nullptr;
};
std::mutex g_mutex;
cudaStream_t g_copy_stream[64] = {};
}  // namespace

extern "C" int paresis_transfer_lane_create(void** lane_out) {
    if (!lane_out) { set_last_error("paresis_transfer_lane_create: null output"); return PARESIS_ERR_ARG; }
    Lane* l = new Lane();
    cudaError_t e = cudaEventCreateWithFlags(&l->ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&l->done, cudaEventDisableTiming);
    if (e != cudaSuccess) { delete l; return check_cuda(e, "cudaEventCreateWithFlags"); }
    *lane_out = l;
    return PARESIS_OK;
}

extern "C" int paresis_transfer_lane_destroy(void* lane) {
    Lane* l = (Lane*)lane;
    if (!l) return PARESIS_OK;
    cudaEventDestroy(l->ready);
    cudaEventDestroy(l->done);
    delete l;
    return PARESIS_OK;
}

// dst_host (pinned) <- src_device, n bytes, once everything queued on `producer` so far has run.
extern "C" int paresis_transfer_d2h(void* lane, void* dst_host, const void* src_device, size_t bytes, paresis_stream producer) {
    Lane* l = (Lane*)lane;
    if (!l || !dst_host || !src_device) { set_last_error("paresis_transfer_d2h: null argument"); return PARESIS_ERR_ARG; }
    int dev = 0;
    PARESIS_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_last_error("paresis_transfer_d2h: device %d out of range", dev); return PARESIS_ERR_ARG; }
    cudaStream_t cs;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        if (!g_copy_stream[dev]) PARESIS_CUDA(cudaStreamCreateWithFlags(&g_copy_stream[dev], cudaStreamNonBlocking));
        cs = g_copy_stream[dev];
    }
    PARESIS_CUDA(cudaEventRecord(l->ready, (cudaStream_t)producer));
    PARESIS_CUDA(cudaStreamWaitEvent(cs, l->ready, 0));
    PARESIS_CUDA(cudaMemcpyAsync(dst_host, src_device, bytes, cudaMemcpyDeviceToHost, cs));
    PARESIS_CUDA(cudaEventRecord(l->done, cs));
    return PARESIS_OK;
}

extern "C" int paresis_transfer_wait(void* lane) {
    Lane* l = (Lane*)lane;
    if (!l) { set_last_error("paresis_transfer_wait: null lane"); return PARESIS_ERR_ARG; }
    PARESIS_CUDA(cudaEventSynchronize(l->done));
    return PARESIS_OK;
}
