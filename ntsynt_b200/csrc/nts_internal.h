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

// Caching device allocator: freed blocks are kept per device and reused.  cudaMalloc / cudaFree are
// synchronising and become much slower once peer access is enabled (NCCL, CUDA IPC), and the hot path
// allocates scratch buffers every step.  Blocks are whole cudaMalloc allocations (never sub-allocated), so a
// pooled pointer can still be exported with cudaIpcGetMemHandle.
cudaError_t pool_alloc(void** p, size_t bytes);
void pool_free(void* p);
void pool_trim(int device);   // give every cached block of a device back to the driver

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
    void release() { if (p) pool_free(p); p = nullptr; n = 0; }
    cudaError_t alloc(size_t count)
    {
        release();
        if (count == 0) count = 1;
        cudaError_t e = pool_alloc(reinterpret_cast<void**>(&p), count * sizeof(T));
        if (e == cudaSuccess) n = count; else p = nullptr;
        return e;
    }
};

}  // namespace nts

namespace nts {
// per-kernel-family device timing (CUDA events on the context's stream), off unless enabled
enum ProfId { PROF_FILL = 0, PROF_BF_INSERT, PROF_BF_COMBINE, PROF_SKETCH, PROF_SKETCH_POST, PROF_JOIN, PROF_SYNTH,
              PROF_POPCOUNT, PROF_BF_REPEAT, PROF_EDGES, PROF_NCCL, PROF_BF_BUILD, PROF_BF_PART1, PROF_BF_PART2, PROF_BF_APPLY,
              PROF_GRAPH, PROF_COUNT };
struct ProfPending { int id; cudaEvent_t e0, e1; double units; };
}  // namespace nts

struct nts_ctx {
    int device = 0;
    bool prof_enabled = false;
    std::vector<nts::ProfPending> prof_pending;
    double prof_ms[nts::PROF_COUNT] = {0};
    double prof_units[nts::PROF_COUNT] = {0};
    uint64_t prof_launches[nts::PROF_COUNT] = {0};
    uint64_t h2d_bytes = 0, d2h_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // side stream of the pipelined Bloom-filter build (created on first use)
    cudaStream_t stream_copy = nullptr;   // H2D stream of nts_genome_upload_async (created on first use)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_side = nullptr;   // orders the side stream against the main one (nts_bf_build_common)
    uint64_t launches = 0;
    uint64_t sketch_escalated = 0;   // dense sub-tiles the sparse sketch kernel handed to the dense one (statistics)
    uint64_t sketch_qall = 0;        // sketches that looked every slot up (low pass rate; statistics)
    uint64_t part_inserts = 0;       // Bloom inserts that took the partitioned path (statistics / tests)
    uint64_t part_overflow_items = 0;   // items those inserts applied through the overflow list
    int sm_count = 148;
    std::map<uint32_t, nts::HashTables*> tables;   // per k, device resident
};

// valid k-mers of a genome for one (k, mask), laid out in valid-index space
namespace nts {
// RAII: times everything launched on ctx->stream during its lifetime under one ProfId
struct ProfScope {
    nts_ctx* ctx; int id; cudaEvent_t e0 = nullptr, e1 = nullptr; double units; uint64_t launches0;
    bool as_one = false;      // count the scope as one unit of work (a pipelined multi-kernel call) instead of its launches
    cudaStream_t st;          // the stream the timed launches go to
    ProfScope(nts_ctx* c, int i, double u = 0, bool one = false, cudaStream_t s = nullptr)
        : ctx(c), id(i), units(u), launches0(c->launches), as_one(one), st(s ? s : c->stream)
    {
        if (ctx->prof_enabled) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
    }
    ~ProfScope()
    {
        ctx->prof_launches[id] += as_one ? 1 : ctx->launches - launches0;
        if (ctx->prof_enabled) { cudaEventRecord(e1, st); ctx->prof_pending.push_back({id, e0, e1, units}); }
    }
};
inline cudaError_t copy_h2d(nts_ctx* ctx, void* dst, const void* src, size_t n)
{
    ctx->h2d_bytes += n;
    return cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, ctx->stream);
}
inline cudaError_t copy_d2h(nts_ctx* ctx, void* dst, const void* src, size_t n)
{
    ctx->d2h_bytes += n;
    return cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, ctx->stream);
}
}  // namespace nts

struct nts_view {
    uint32_t k = 0;
    nts::DevBuf<uint64_t> seg_v, seg_base;
    uint32_t n_seg = 0;
    std::vector<uint64_t> contig_v;   // [n_contigs + 1] valid-index range of each contig
    uint64_t total_valid = 0;
    std::vector<uint64_t> h_seg_v, h_seg_base;   // host copies of the island table (a few thousand entries)
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
    cudaEvent_t ready = nullptr;             // async upload: recorded after the H2D copy on ctx->stream_copy
    bool ready_waited = true;                // ctx->stream already ordered after `ready`
    // the async copy goes in growing chunks, an event after each: words [0, chunk_end_word[i]) are on the device once
    // chunk_ev[i] has fired (the last one is `ready`), so the first consumer can start on the head of the genome
    std::vector<cudaEvent_t> chunk_ev;
    std::vector<uint64_t> chunk_end_word;
};

// one stage of a Bloom insert that starts while the genome is still being uploaded: after `ev`, the k-mers with valid
// index < v_end have all their bases on the device
struct UploadStage { cudaEvent_t ev; uint64_t v_end; };

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
