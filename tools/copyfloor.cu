// copyfloor.cu -- pure-CUDA host<->device copy floor (a measuring tool, NOT part of libedge_b200.so).
//
// bench.py's end-to-end leg moves one fp32 batch host -> device and one device -> host per step.  This tool measures what
// the box can deliver for exactly that traffic with nothing else in the way: pinned buffers allocated (and first-touched)
// by the calling thread AFTER it has been bound to the GPU's NUMA node, one cudaMemcpyAsync per chunk, H2D and D2H on two
// dedicated streams, timed with CUDA events.  Every rank calls it at the same time (bench.py puts a barrier in front), so
// at N ranks it measures the N-way contended floor of the host memory / PCIe fabric.
//
//   mode 0: H2D only    mode 1: D2H only    mode 2: both directions concurrently (what a pipelined step does)
//   returns 0 and *ms_per_iter (mean over `iters` after 2 warm-up iterations), or a cudaError_t.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#define CF_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = (int)e_; goto done; } } while (0)

extern "C" int cf_measure(int device, size_t bytes, int chunks, int iters, int mode, int use_host_register, double* ms_per_iter) {
    int rc = 0;
    void *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
    void *raw_in = nullptr, *raw_out = nullptr;
    cudaStream_t s_up = nullptr, s_dn = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e_dn = nullptr;
    if (chunks < 1) chunks = 1;
    const size_t chunk = (bytes / chunks + 255) & ~(size_t)255;
    CF_CHECK(cudaSetDevice(device));
    if (use_host_register) {            // malloc + first touch on this (NUMA-bound) thread, then pin in place
        if (posix_memalign(&raw_in, 4096, bytes) || posix_memalign(&raw_out, 4096, bytes)) { rc = -1; goto done; }
        memset(raw_in, 1, bytes); memset(raw_out, 0, bytes);
        CF_CHECK(cudaHostRegister(raw_in, bytes, cudaHostRegisterDefault));
        CF_CHECK(cudaHostRegister(raw_out, bytes, cudaHostRegisterDefault));
        h_in = raw_in; h_out = raw_out;
    } else {
        CF_CHECK(cudaHostAlloc(&h_in, bytes, cudaHostAllocDefault));
        CF_CHECK(cudaHostAlloc(&h_out, bytes, cudaHostAllocDefault));
        memset(h_in, 1, bytes); memset(h_out, 0, bytes);
    }
    CF_CHECK(cudaMalloc(&d_in, bytes));
    CF_CHECK(cudaMalloc(&d_out, bytes));
    CF_CHECK(cudaMemset(d_out, 0, bytes));
    CF_CHECK(cudaStreamCreateWithFlags(&s_up, cudaStreamNonBlocking));
    CF_CHECK(cudaStreamCreateWithFlags(&s_dn, cudaStreamNonBlocking));
    CF_CHECK(cudaEventCreate(&e0)); CF_CHECK(cudaEventCreate(&e1)); CF_CHECK(cudaEventCreate(&e_dn));
    for (int it = -2; it < iters; ++it) {
        if (it == 0) {
            CF_CHECK(cudaStreamSynchronize(s_up)); CF_CHECK(cudaStreamSynchronize(s_dn));
            CF_CHECK(cudaEventRecord(e0, s_up));
            CF_CHECK(cudaStreamWaitEvent(s_dn, e0, 0));
        }
        for (size_t off = 0; off < bytes; off += chunk) {
            const size_t n = (bytes - off < chunk) ? bytes - off : chunk;
            if (mode == 0 || mode == 2) CF_CHECK(cudaMemcpyAsync((char*)d_in + off, (char*)h_in + off, n, cudaMemcpyHostToDevice, s_up));
            if (mode == 1 || mode == 2) CF_CHECK(cudaMemcpyAsync((char*)h_out + off, (char*)d_out + off, n, cudaMemcpyDeviceToHost, s_dn));
        }
    }
    CF_CHECK(cudaEventRecord(e_dn, s_dn));
    CF_CHECK(cudaStreamWaitEvent(s_up, e_dn, 0));
    CF_CHECK(cudaEventRecord(e1, s_up));
    CF_CHECK(cudaEventSynchronize(e1));
    {
        float ms = 0.0f;
        CF_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        *ms_per_iter = (double)ms / (iters > 0 ? iters : 1);
    }
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (e_dn) cudaEventDestroy(e_dn);
    if (s_up) cudaStreamDestroy(s_up);
    if (s_dn) cudaStreamDestroy(s_dn);
    if (d_in) cudaFree(d_in);
    if (d_out) cudaFree(d_out);
    if (use_host_register) {
        if (raw_in) { cudaHostUnregister(raw_in); free(raw_in); }
        if (raw_out) { cudaHostUnregister(raw_out); free(raw_out); }
    } else {
        if (h_in) cudaFreeHost(h_in);
        if (h_out) cudaFreeHost(h_out);
    }
    return rc;
}
