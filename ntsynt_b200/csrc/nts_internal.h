// nts_internal.h -- host-side object definitions behind the opaque C-ABI handles.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/ntsynt_b200.h"
#include "nts_device.cuh"

namespace nts {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define NTS_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (call);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return ::nts::fail(NTS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));        \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; return *this; }
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    cudaError_t alloc(size_t count)
    {
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
        if (e == cudaSuccess) n = count; else p = nullptr;
        return e;
    }
};

}  // namespace nts

struct nts_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint64_t launches = 0;
    int sm_count = 148;
    std::map<uint32_t, nts::HashTables*> tables;   // per k, device resident
};

// valid k-mers of a genome for one (k, mask), laid out in valid-index space
struct nts_view {
    uint32_t k = 0;
    nts::DevBuf<uint64_t> seg_v, seg_base;
    uint32_t n_seg = 0;
    std::vector<uint64_t> contig_v;   // [n_contigs + 1] valid-index range of each contig
    uint64_t total_valid = 0;
};

struct nts_genome {
    nts_ctx* ctx = nullptr;
    uint32_t n_contigs = 0;
    std::vector<uint64_t> contig_len;
    std::vector<uint64_t> contig_word_off;   // word offset of each contig in `packed`
    uint64_t n_words = 0;
    uint64_t total_bases = 0;
    nts::DevBuf<uint64_t> packed;
    std::vector<uint64_t> nrun_off, nrun_start, nrun_len;   // host copy (tiny)
    std::map<uint32_t, nts_view*> views;     // unmasked views per k
};

struct nts_bf {
    nts_ctx* ctx = nullptr;
    uint64_t bytes = 0;          // logical size (multiple of 8)
    uint64_t alloc_bytes = 0;    // padded to a multiple of 16 (padding stays zero)
    nts::DevBuf<uint32_t> words;
};

struct nts_mxs {
    nts_ctx* ctx = nullptr;
    uint64_t count = 0;
    uint32_t n_contigs = 0;
    nts::DevBuf<uint64_t> h1;
    nts::DevBuf<uint32_t> pos, contig;
};
