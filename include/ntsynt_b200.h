/*
 * ntsynt_b200.h -- C-ABI of libntsynt_b200.so: the B200-native (sm_100a) replacement for
 * ntSynt's minimizer-sketch -> common-Bloom-filter -> minimizer-graph hot path.
 *
 * The reference (bcgsc/ntSynt v1.0.4, paths below are relative to its tree) has no
 * in-process FFI for this path: its seam is three executables plus the files between them
 * (bin/ntsynt_run_pipeline.smk:48-103).  Each entry point below names the reference call
 * site it replaces; INTEGRATION.md shows the ctypes binding and the drop-in executables.
 *
 * Conventions: every function returns 0 on success and a negative nts_status on failure;
 * nts_last_error() gives the message of the last failure on the calling thread.  All
 * pointers are HOST pointers owned by the caller unless the name says `_dev`.  One CUDA
 * stream per context; calls are synchronous unless named `_async`.  A context is
 * thread-compatible (distinct contexts on distinct threads), not thread-safe.
 * There is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef NTSYNT_B200_H
#define NTSYNT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    NTS_OK = 0,
    NTS_ERR_CUDA = -1,     /* a CUDA runtime call failed */
    NTS_ERR_ARG = -2,      /* invalid argument */
    NTS_ERR_NOMEM = -3,    /* host or device allocation failed */
    NTS_ERR_OVERFLOW = -4, /* an output did not fit its capacity even after the retry */
    NTS_ERR_NCCL = -5,     /* NCCL not loadable or a NCCL call failed */
    NTS_ERR_STATE = -6     /* call sequence error */
} nts_status;

typedef struct nts_ctx nts_ctx;       /* one per GPU / host thread */
typedef struct nts_genome nts_genome; /* 2-bit packed contigs + N runs, resident in HBM */
typedef struct nts_bf nts_bf;         /* Bloom filter bit array in HBM (1 hash function) */
typedef struct nts_mxs nts_mxs;       /* minimizer table (h1, pos, contig) in HBM */

/* -------------------------------------------------------------------------------- library */
const char* nts_version(void);
const char* nts_last_error(void);
int nts_device_count(int* out);

/* -------------------------------------------------------------------------------- context */
int nts_ctx_create(int device, nts_ctx** out);
void nts_ctx_destroy(nts_ctx* ctx);
int nts_ctx_sync(nts_ctx* ctx);
/* CUDA-event timer on the context's stream (used by bench.py for device-side timing). */
int nts_timer_start(nts_ctx* ctx);
int nts_timer_stop(nts_ctx* ctx, float* ms_out);
/* number of kernel launches issued by this library on this context since creation */
uint64_t nts_launch_count(const nts_ctx* ctx);
/* statistics: dense sub-tiles the sparse sketch kernel handed to the dense selector (unresolved windows) since creation */
uint64_t nts_sketch_escalated(const nts_ctx* ctx);
/* sketches that looked EVERY slot up because the filter passes < 10 % of the k-mers (statistics) */
uint64_t nts_sketch_queried_all(const nts_ctx* ctx);
/* statistics: Bloom inserts that took the partitioned path (csrc/nts_part.cuh), and the items those inserts applied
 * through their overflow lists (heavy-hitter k-mers), since creation */
uint64_t nts_part_inserts(const nts_ctx* ctx);
uint64_t nts_part_overflow_items(const nts_ctx* ctx);
/* free / total device memory in bytes */
int nts_mem_info(nts_ctx* ctx, uint64_t* free_b, uint64_t* total_b);
/* Per-kernel-family device timing with CUDA events on the context's stream (bench.py's roofline
 * leg).  Families: nts_prof_name(0..nts_prof_count()-1).  units = k-mers / bytes / items the
 * launches processed (family specific). */
int nts_prof_enable(nts_ctx* ctx, int on);
int nts_prof_reset(nts_ctx* ctx); /* also zeroes the transfer counters */
int nts_prof_count(void);
const char* nts_prof_name(int id);
int nts_prof_get(nts_ctx* ctx, int id, double* ms, double* units, uint64_t* launches);
/* bytes moved host->device / device->host by this library on this context */
int nts_xfer_bytes(const nts_ctx* ctx, uint64_t* h2d, uint64_t* d2h);
/* page-locked host memory for genome uploads */
int nts_host_alloc(uint64_t bytes, void** out);
void nts_host_free(void* p);

/* -------------------------------------------------------------------------------- ingest
 * Replaces btllib::SeqReader as used at src/ntsynt_make_common_bf.cpp:32-36,125,143 and
 * inside indexlr.  Packed layout: 2 bits per base (A=0,C=1,G=2,T=3; anything else is stored
 * as 0 and recorded as an N run), 32 bases per little-endian uint64 word, base i of a contig
 * at bits [2*(i%32), 2*(i%32)+1] of word i/32.  Lower-case acgt pack like upper case.
 */
/* number of uint64 words a contig of n bases occupies (rounded up to a 16-byte multiple) */
uint64_t nts_packed_words(uint64_t n_bases);
/* Pack one ASCII sequence.  nrun_start/nrun_len receive maximal runs of non-ACGT characters
 * (capacity nrun_cap each); *n_nruns gets the number found (may exceed nrun_cap: call again
 * with larger buffers).  words_out must hold nts_packed_words(n) words. */
int nts_pack_ascii(const char* seq, uint64_t n, uint64_t* words_out, uint64_t* nrun_start, uint64_t* nrun_len,
                   uint64_t nrun_cap, uint64_t* n_nruns);
/* Inverse (for --seq output and tests): bases [start, start+n) of a packed contig -> ASCII ACGT. */
int nts_unpack_ascii(const uint64_t* words, uint64_t start, uint64_t n, char* out);

/* Upload a genome.  contig c occupies nts_packed_words(contig_len[c]) words starting at word
 * contig_word_off[c] of `words` (caller lays contigs out back to back in that order).
 * N runs of contig c are entries [nrun_off[c], nrun_off[c+1]) of nrun_start/nrun_len (sorted,
 * contig-local coordinates). */
int nts_genome_upload(nts_ctx* ctx, uint32_t n_contigs, const uint64_t* contig_len, const uint64_t* contig_word_off,
                      const uint64_t* words, uint64_t n_words, const uint64_t* nrun_off, const uint64_t* nrun_start,
                      const uint64_t* nrun_len, nts_genome** out);
/* Same, but the host->device copy is queued on the context's copy stream and the call returns at once; every entry
 * point that reads the genome orders itself after the copy.  `words` must stay valid (and should be page-locked,
 * nts_host_alloc) until the first such call has completed.  Lets the upload of genome i+1 overlap the Bloom insert
 * of genome i (the reference re-reads every FASTA per stage instead: src/ntsynt_make_common_bf.cpp:125,143). */
int nts_genome_upload_async(nts_ctx* ctx, uint32_t n_contigs, const uint64_t* contig_len, const uint64_t* contig_word_off,
                      const uint64_t* words, uint64_t n_words, const uint64_t* nrun_off, const uint64_t* nrun_start,
                      const uint64_t* nrun_len, nts_genome** out);
void nts_genome_destroy(nts_genome* g);
/* total bases, Ns included (the `n` of approximate_bf_size, cpp:34-36) */
uint64_t nts_genome_size(const nts_genome* g);
uint32_t nts_genome_contigs(const nts_genome* g);
/* copy the packed words of one contig back to the host (nts_packed_words(len) words) */
int nts_genome_download_contig(nts_genome* g, uint32_t contig, uint64_t* words_out);
/* N runs of the whole genome, same layout as nts_genome_upload's arguments */
int nts_genome_nruns(nts_genome* g, uint64_t* nrun_off /*[n_contigs+1]*/, uint64_t* nrun_start, uint64_t* nrun_len,
                     uint64_t cap, uint64_t* n_out);

/* Synthetic genome materialised on the device from a segment table (bench.py workloads).
 * Output base j of contig c is addressed by the segment covering j: a copy of an ancestor base
 * (strand +1) or of its complement walking backwards (strand -1), substituted with probability
 * sub_rate; a random inserted base (anc_contig == -1); or N (anc_contig == -2).  The ancestor is a
 * pure function of (anc_seed, contig, position): i.i.d. bases with P(A)=P(T)=0.295,
 * P(C)=P(G)=0.205 overlaid with copies of n_repeat_fam repeat families (300 bp and 6 kbp), one
 * copy per 8 kbp slot with probability repeat_slot_prob, each copy mutated at 10 %.
 * segs: [n_seg] sorted by (dst_contig, dst_start), tiling every contig without gaps. */
typedef struct {
    uint32_t dst_contig;
    int32_t anc_contig;   /* >=0 ancestor contig; -1 random insert; -2 N run */
    uint64_t dst_start;   /* contig-local start in the new genome */
    uint64_t anc_start;   /* ancestor coordinate of the segment's FIRST output base */
    uint64_t len;
    int32_t strand;       /* +1 forward, -1 reverse complement (ancestor coordinate then decreases) */
    uint32_t pad;
} nts_synth_seg;
int nts_genome_synthesize(nts_ctx* ctx, uint32_t n_contigs, const uint64_t* contig_len, const nts_synth_seg* segs,
                          uint64_t n_seg, uint64_t anc_seed, uint64_t genome_seed, double sub_rate,
                          uint32_t n_repeat_fam, double repeat_slot_prob, nts_genome** out);
/* host evaluation of the ancestor formula (tests) */
int nts_synth_ancestor_base(uint64_t anc_seed, uint32_t n_repeat_fam, double repeat_slot_prob, uint32_t contig,
                            uint64_t pos);

/* -------------------------------------------------------------------------------- Bloom filter
 * Replaces btllib::KmerBloomFilter(bytes, 1, k) as used by src/ntsynt_make_common_bf.cpp. */
/* a1: approximate_bf_size (cpp:28-40) + btllib's round-up to a multiple of 8 bytes */
uint64_t nts_bf_bytes(int64_t genome_size, double fpr);
int nts_bf_create(nts_ctx* ctx, uint64_t bytes, nts_bf** out); /* zero-filled */
void nts_bf_destroy(nts_bf* bf);
uint64_t nts_bf_size_bytes(const nts_bf* bf);
int nts_bf_clear(nts_bf* bf);
/* a2: bf->insert(record.seq) for every record (cpp:128-131): sets bit ntHash2_h0(kmer) mod m
 * for every all-ACGT k-mer (kernels i + iii-a). */
int nts_bf_insert_genome(nts_bf* bf, const nts_genome* g, uint32_t k);
/* bf = bits(g): nts_bf_clear + nts_bf_insert_genome without the zero-fill pass (the first level of cpp:122-131) */
int nts_bf_set_genome(nts_bf* bf, const nts_genome* g, uint32_t k);
/* a3: the cascade of cpp:136-160 with one hash function == dst &= src (kernel iii-b) */
int nts_bf_and(nts_bf* dst, const nts_bf* src);
int nts_bf_or(nts_bf* dst, const nts_bf* src);
/* The whole of src/ntsynt_make_common_bf.cpp:107-160 in one call:
 * common = AND over i of bits(genomes[i]) (genomes in the caller's sorted-path order; `level` is scratch of the same
 * size, may be NULL when n == 1; neither needs to be zeroed).  Genome 0 is written as a whole filter, every further
 * genome as next = current & bits(genome) into the other filter (the cascade level of cpp:136-160), so there is no
 * zero-fill and no separate AND pass; the two filters may swap their device storage.  Result is bit-identical to
 * nts_bf_insert_genome + nts_bf_and. */
int nts_bf_build_common(nts_bf* common, nts_bf* level, const nts_genome* const* genomes, uint32_t n, uint32_t k);
/* the same without the last AND pass (one read + one write of both 14.8 GB arrays): on return, when *level_is_apart = 1,
 * `common` = AND over genomes 0 .. n-2 and `level` = bits(genome n-1) -- their AND is the common filter of
 * src/ntsynt_make_common_bf.cpp:136-160; nts_sketch2 takes the pair, nts_bf_and(common, level) makes it one filter when
 * one is needed (saving it, merging across GPUs).  *level_is_apart = 0: `common` is complete, `level` is scratch. */
int nts_bf_build_common_lazy(nts_bf* common, nts_bf* level, const nts_genome* const* genomes, uint32_t n, uint32_t k,
                             int* level_is_apart);
/* repeat filter, bin/ntsynt_make_repeat_bfs.py:56-69: rep |= k-mers seen >= 2x in g (kernel iii-d).
 * `scratch` is a per-genome filter of the same size that the call clears and uses. */
int nts_bf_insert_repeats(nts_bf* rep, nts_bf* scratch, const nts_genome* g, uint32_t k);
int nts_bf_popcount(nts_bf* bf, uint64_t* bits_set); /* for the "Bloom filter FPR" log lines */
int nts_bf_download(nts_bf* bf, uint8_t* bytes_out);
int nts_bf_upload(nts_bf* bf, const uint8_t* bytes_in);
/* -------------------------------------------------------------------------------- sketch
 * Replaces `indexlr -k K -w W --long --pos [-s common.bf] [-r repeat.bf]`
 * (bin/ntsynt_run_pipeline.smk:83-85; subprojects/ntJoin/bin/ntjoin_utils.py:195-202).
 * mask_*: optional extra N intervals (the synteny-block masks of
 * bin/ntsynt_synteny.py:134-157), contig c owning entries [mask_off[c], mask_off[c+1]),
 * sorted, half-open [start, end), contig-local.  Pass mask_off = NULL for none.
 * Output order: by contig, then by position (== indexlr's emission order). */
int nts_sketch(nts_ctx* ctx, const nts_genome* g, const nts_bf* common /*nullable*/, const nts_bf* repeat /*nullable*/,
               uint32_t k, uint32_t w, const uint64_t* mask_off, const uint64_t* mask_start,
               const uint64_t* mask_end, nts_mxs** out);
/* nts_sketch with the common filter given as two filters of equal size whose AND it is (common2 may be null) */
int nts_sketch2(nts_ctx* ctx, const nts_genome* g, const nts_bf* common /*nullable*/, const nts_bf* common2 /*nullable*/,
                const nts_bf* repeat /*nullable*/, uint32_t k, uint32_t w, const uint64_t* mask_off, const uint64_t* mask_start,
                const uint64_t* mask_end, nts_mxs** out);
void nts_mxs_destroy(nts_mxs* m);
uint64_t nts_mxs_count(const nts_mxs* m);
/* h1 = second ntHash2 hash (what indexlr prints), pos = 0-based k-mer start, contig index */
int nts_mxs_download(nts_mxs* m, uint64_t* h1, uint32_t* pos, uint32_t* contig);
/* build a device table from host arrays (bin/ntsynt_run.py reads indexlr TSVs: ntjoin_utils.py:167-193) */
int nts_mxs_upload(nts_ctx* ctx, uint64_t n, const uint64_t* h1, const uint32_t* pos, const uint32_t* contig,
                   nts_mxs** out);
/* test hook: canonical h0 and validity of every k-mer start of one contig (kernel i alone) */
int nts_hash_contig(nts_ctx* ctx, const nts_genome* g, uint32_t contig, uint32_t k, uint64_t* h0_out,
                    uint8_t* valid_out);

/* -------------------------------------------------------------------------------- multi-GPU
 * The one inter-GPU exchange of the path (the reference is a single process: cpp:136-160 cascades
 * genome after genome).  Every rank holds a filter; after nts_bf_allreduce_and each rank's filter is
 * the AND over all ranks (OR for nts_bf_allreduce_or: contig-sharded partial filters of one genome).
 * Implementation: bits expanded to 2/4/8-bit counters packed in uint32, ONE ncclAllReduce(sum) per
 * 128 MB chunk (double buffered), thresholded back to bits.  NCCL is dlopen()ed on first use.
 * The 128-byte id comes from rank 0 (nts_nccl_unique_id) over any host side channel. */
typedef struct nts_comm nts_comm;
int nts_nccl_unique_id(uint8_t id_out[128]);
int nts_nccl_init(nts_ctx* ctx, const uint8_t id[128], int rank, int world, nts_comm** out);
void nts_nccl_destroy(nts_comm* comm);
/* stream-ordered barrier (a one-word all-reduce on the context's stream; no host synchronisation) */
int nts_nccl_barrier(nts_comm* c);
int nts_nccl_world(const nts_comm* comm);
int nts_nccl_rank(const nts_comm* comm);
int nts_bf_allreduce_and(nts_comm* comm, nts_bf* bf);
int nts_bf_allreduce_or(nts_comm* comm, nts_bf* bf);
/* Peer-memory merge (B200-native alternative, same result): every rank exports its filter with CUDA IPC
 * (nts_bf_ipc_handle), maps the peers' filters (nts_p2p_open, handles in rank order), then
 *   barrier; nts_p2p_reduce_scatter(op: 0 AND, 1 OR); barrier; nts_p2p_all_gather; barrier
 * The reduce kernel reads slice `rank` of every peer's array straight from peer HBM over NVLink and
 * combines it in registers; wire volume per rank 2*(P-1)/P * filter bytes. */
typedef struct nts_p2p nts_p2p;
int nts_bf_ipc_handle(nts_bf* bf, uint8_t handle_out[64]);
int nts_p2p_open(nts_bf* mine, const uint8_t* handles, int rank, int world, nts_p2p** out);
void nts_p2p_close(nts_p2p* p);
int nts_p2p_reduce_scatter(nts_p2p* p, int op);
int nts_p2p_all_gather(nts_p2p* p);
/* Contig-sharded ownership (G < number of GPUs): every rank holds, per genome, the bits of its own
 * contigs.  sets[g] maps genome g's partial filters of all ranks, `out` maps the common filter:
 *   barrier; nts_p2p_reduce_and_of_or(sets, G, out); barrier; nts_p2p_all_gather(out); barrier
 * computes common = AND over genomes of (OR over ranks) -- src/ntsynt_make_common_bf.cpp:136-160 and the cross-GPU
 * merge in one kernel that reads peer HBM; up to 8 genomes. */
int nts_p2p_reduce_and_of_or(nts_p2p* const* sets, uint32_t n_sets, nts_p2p* out);

/* ---- hash-range OWNED build of the common filter (multi-GPU; replaces building whole per-GPU filters and merging them,
 * src/ntsynt_make_common_bf.cpp:122-160 across GPUs).  Every GPU bins its k-mers with one agreed plan (pass 1 of the
 * partitioned insert); the GPU that owns a byte range of the filter -- the slices of nts_p2p_all_gather -- applies the
 * buckets of every GPU to it, reading them over NVLink peer memory (4 bytes per k-mer on the wire), ORs per genome, ANDs
 * across genomes on its slice (nts_bf_range_op), and nts_p2p_all_gather distributes the finished filter. */
typedef struct nts_binpeer nts_binpeer;
/* plan + scratch of `slot` for filters of sized_like's size and up to plan_valid k-mers per binning pass (same values on
 * every rank); the scratch is what nts_bin_ipc_handles exports */
int nts_bin_prepare(nts_bf* sized_like, int slot, uint64_t plan_valid);
/* pass 1 only, asynchronous: the genome's k-mers binned by filter region into the slot's buckets */
int nts_bin_genome(nts_bf* sized_like, const nts_genome* g, uint32_t k, int slot);
/* k-mers that did not fit their bucket since nts_bin_prepare (heavy hitters; they were NOT applied: redo with the merge) */
int nts_bin_overflow(nts_ctx* ctx, int slot, uint64_t* n);
int nts_bin_ipc_handles(nts_ctx* ctx, int slot, uint8_t items_handle[64], uint8_t cursor_handle[64]);
int nts_binpeer_open(nts_ctx* ctx, const uint8_t* items_handles /* world x 64 */, const uint8_t* cursor_handles, int rank, int world,
                     int slot, nts_binpeer** out);
void nts_binpeer_close(nts_binpeer* p);
/* bf[off16 .. off16 + n16) |= bits of what rank `source` binned (units of 16 bytes; asynchronous) */
int nts_bf_apply_owned(nts_bf* bf, nts_binpeer* p, int source, uint64_t off16, uint64_t n16);
/* dst[range] op= src[range]: op 0 AND, 1 OR, 2 COPY (src null: zero); asynchronous */
int nts_bf_range_op(nts_bf* dst, const nts_bf* src, uint64_t off16, uint64_t n16, int op);
/* the slice rank s owns in nts_p2p_reduce_scatter / nts_p2p_all_gather */
int nts_p2p_slice(const nts_p2p* p, int s, uint64_t* off16, uint64_t* n16);
/* all-gather of minimizer tables (ncclAllGather over padded columns): out[r] = rank r's table, as a
 * new nts_mxs on this rank; counts[world] must hold every rank's table size. */
int nts_mxs_allgather(nts_comm* comm, const nts_mxs* mine, const uint64_t* counts, nts_mxs** out);

/* -------------------------------------------------------------------------------- graph (kernel iv)
 * Replaces ntjoin_utils.read_minimizers' duplicate removal (subprojects/ntJoin/bin/ntjoin_utils.py:182-192),
 * filter_minimizers (:152-165), build_graph's adjacency edges and weights (:97-113,132-135) and,
 * for edges supported by every assembly, filter_graph_global + find_paths
 * (subprojects/ntJoin/bin/ntjoin.py:78-87,114-136).
 * Input: one minimizer table per assembly, in the caller's assembly order (= the reference's
 * reverse-sorted FILES order, bin/ntsynt_synteny.py:34).  A minimizer becomes a vertex iff it occurs
 * exactly once in every assembly.  Vertices are numbered by their rank in assembly `order_asm`'s
 * filtered list (contig order, then position), so an edge supported by all assemblies is always
 * (i, i+1) and maximal chains are runs of the `link` bitmap. */
typedef struct nts_graph nts_graph;
int nts_graph_build(nts_ctx* ctx, nts_mxs* const* tables, uint32_t n_asm, uint32_t order_asm, nts_graph** out);
void nts_graph_destroy(nts_graph* g);
uint64_t nts_graph_vertices(const nts_graph* g);
/* Vertex table (any pointer may be NULL): h1[V]; pos / contig / rank are [n_asm*V], assembly-major,
 * rank = index of the vertex in that assembly's filtered list (all contigs concatenated);
 * link[V]: 1 iff the edge (i, i+1) has weight n_asm; degree[V]: number of distinct neighbours. */
int nts_graph_download_vertices(nts_graph* g, uint64_t* h1, uint32_t* pos, uint32_t* contig, uint32_t* rank,
                                uint8_t* link, uint8_t* degree);
/* Per-pair arrays for (i, i+1) in vertex order, consumed by the host as prefix sums over chains:
 * inv[n_asm*V] = vertex at each rank (inverse of rank); incmask/decmask[V]: bit a set iff assembly a's
 * position increases / decreases from vertex i to i+1 (orientation rule, bin/synteny_block.py:48-65);
 * spread[V] = max_a |dpos| - min_a |dpos| (indel test, bin/ntsynt_synteny.py:364-368,399). */
int nts_graph_download_links(nts_graph* g, uint32_t* inv, uint32_t* incmask, uint32_t* decmask, uint32_t* spread);
/* prefix sums of the direction bits (device scans): ci, cd are [n_asm*(V+1)], assembly-major;
 * ci[a*(V+1)+i] = number of pairs (j, j+1), j < i, whose position increases in assembly a (cd: decreases) */
int nts_graph_download_cums(nts_graph* g, uint32_t* ci, uint32_t* cd);
/* Host-ready columns for the graph stage, with room to grow: `cap` >= V is the row pitch (in elements) of
 * pos64 / ctg32 and the length of h1 / nbr / conn; only the first V entries of each row are written.
 * pos64 = positions widened to int64; nbr[v] = (left, right) neighbour of v in the weight-filtered graph
 * (subprojects/ntJoin/bin/ntjoin.py:78-87) or -1; conn[i] = 1 iff edge (i, i+1) has full weight. */
int nts_graph_download_host_arrays(nts_graph* g, uint64_t cap, uint64_t* h1, long long* pos64, int32_t* ctg32, int32_t* nbr,
                                   uint8_t* conn);
/* Sparse views of the vertex table, each sorted ascending: breaks = i < V-1 without a full-weight edge (i, i+1);
 * deg3 = vertices with exactly 3 distinct neighbours (candidates of run_graph_simplification,
 * bin/ntsynt_synteny.py:566-590); big = i whose pair (i, i+1) spreads by more than bp (check_for_indels,
 * bin/ntsynt_synteny.py:364-409).  Pass NULL pointers to get the counts first. */
int nts_graph_sparse_lists(nts_graph* g, uint32_t bp, uint32_t* breaks, uint32_t* deg3, uint32_t* big, uint64_t counts[3]);
/* vertex id of each h1 (0xFFFFFFFF if none) via the join table kept on the device
 * (replaces `mx in mx_info` / graph.vs.find(name) lookups of the refinement rounds) */
int nts_graph_lookup(nts_graph* g, const uint64_t* h1, uint64_t n, uint32_t* vid_out);
/* Edge table: distinct unordered adjacencies in build_graph's first-insertion order (assembly
 * order, then list order); (u, v) in the orientation of the first insertion; support = bitmask of
 * assemblies (weight = popcount, all assembly weights are 1: bin/ntsynt_synteny.py:32). */
int nts_graph_edges(nts_graph* g, uint64_t* n_edges);
int nts_graph_download_edges(nts_graph* g, uint32_t* u, uint32_t* v, uint32_t* support);
/* --filter Filter of the graph stage (bin/ntsynt_synteny.py:601-609; read_minimizers(tsv, repeat_bf),
 * subprojects/ntJoin/bin/ntjoin_utils.py:182): out = the entries of `m` (a table of genome g) whose k-mer is not in bf */
int nts_mxs_drop_in_bf(nts_ctx* ctx, const nts_mxs* m, const nts_genome* g, const nts_bf* bf, uint32_t k, nts_mxs** out);
/* row range [off[c], off[c+1]) of every contig of a table (off has n_contigs + 1 entries); a new table made of row
 * ranges of other tables in the given order -- used to put the tables of a contig-sharded sketch back in contig order */
int nts_mxs_contig_offsets(nts_mxs* m, uint32_t n_contigs, uint64_t* off);
int nts_mxs_concat(nts_ctx* ctx, nts_mxs* const* parts, const uint64_t* src_off, const uint64_t* cnt, uint64_t n_parts,
                   nts_mxs** out);
/* ---- lean form of the graph stage: the O(V) columns stay on the device, the host asks for what it walks ---------
 * nts_graph_gather: columns at the n vertex ids idx (what: 0 h1 -> u64[n]; 1 pos -> i64[n_asm x n]; 2 contig ->
 * i32[n_asm x n]; 3 rank -> u32[n_asm x n]; 4 inv = vertex at rank idx -> u32[n_asm x n]); ids outside [0, V) give 0. */
int nts_graph_gather(nts_graph* g, int what, const int64_t* idx, uint64_t n, void* out);
/* up / down [n_asm x n] i64: pairs (j, j+1), lo[i] <= j < hi[i] (clamped to V), whose position increases / decreases
 * in each assembly -- the orientation tallies of bin/synteny_block.py:48-65 from device prefix sums */
int nts_graph_range_sums(nts_graph* g, const int64_t* lo, const int64_t* hi, uint64_t n, long long* up, long long* down);
/* left / right / rank [n x n_asm] i64 of the vertices cand: neighbour on the same contig line in every assembly's
 * filtered list or -1 (the neighbourhoods run_graph_simplification inspects, bin/ntsynt_synteny.py:548-590) */
int nts_graph_neigh(nts_graph* g, const int64_t* cand, uint64_t n, long long* left, long long* right, long long* rk);
/* nbr[cap x 2] i32 and conn[cap] u8 of the weight-filtered graph alone (see nts_graph_download_host_arrays) */
int nts_graph_download_links_nbr(nts_graph* g, uint64_t cap, int32_t* nbr, uint8_t* conn);
/* Chain extraction on the device (find_paths, subprojects/ntJoin/bin/ntjoin.py:114-136, for the (i, i+1) part of the
 * graph): ascending (starts, ends) of the maximal runs of >= 2 base vertices joined by full-weight edges.  Call with
 * NULL outputs for the count, then with arrays of that length.  nts_graph_set_links pushes the host's edits of pairs
 * (i, i+1) (val 1 = edge present) back to the device's link bitmap first. */
int nts_graph_set_links(nts_graph* g, const int64_t* idx, const uint8_t* val, uint64_t n);
int nts_graph_runs(nts_graph* g, int64_t* starts, int64_t* ends, uint64_t* n_runs);
/* One refinement round's minimizer filtering on the device, for the new (masked, smaller-w) tables of all assemblies:
 *   read_minimizers' duplicate removal (subprojects/ntJoin/bin/ntjoin_utils.py:182-192), find_mx_in_blocks +
 *   filter_minimizers_synteny_blocks (bin/ntsynt_synteny.py:205-280: drop the blocks' internal minimizers and the ones
 *   inside a block interval, start a new sub-list where the span from the previous kept minimizer overlaps a block)
 *   and filter_minimizers (ntjoin_utils.py:152-165: keep the keys that survive in every assembly).
 * Host inputs (small): the blocks' vertex segments (seg_lo / seg_hi ascending, disjoint), their terminal vertices
 * (ascending), the vertices added after round 0 (x_key ascending with x_vid), and per assembly a the block intervals
 * iv_start[iv_off[a] .. iv_off[a+1]) ascending with iv_maxend their running maximum end (coordinate = contig << 40 | pos).
 * Outputs: n_raw[a] = entries of assembly a after duplicate removal; the surviving (h1, pos, ctg, sub-list id) of
 * assembly a at [out_off[a], out_off[a+1]) of the out_* arrays (out_cap entries in all; if out_off[n_asm] > out_cap only
 * out_off is valid -- call again with larger arrays). */
int nts_graph_refine_filter(nts_graph* g, nts_mxs* const* tables, const uint32_t* seg_lo, const uint32_t* seg_hi, uint32_t n_seg,
                            const uint32_t* term, uint32_t n_term, const uint64_t* x_key, const uint32_t* x_vid, uint32_t n_x,
                            const long long* iv_start, const long long* iv_maxend, const uint64_t* iv_off, uint64_t* n_raw,
                            uint64_t* out_off, uint64_t* out_h1, uint32_t* out_pos, uint32_t* out_ctg, uint32_t* out_sub,
                            uint64_t out_cap);
/* number of pairs (i, i+1) whose |dpos| spread exceeds bp */
int nts_graph_big_count(nts_graph* g, uint32_t bp, uint64_t* n_big);
/* Kernel (iv-c), collinear-path extraction on the device, for the n_runs runs [starts[i], ends[i]] of base vertices
 * joined by (i, i+1) full-weight edges whose vertices still hold their round-0 positions: find_paths
 * (subprojects/ntJoin/bin/ntjoin.py:89-136), find_synteny_blocks + orientation (bin/ntsynt_synteny.py:66-106,
 * bin/synteny_block.py:48-65; m_pct = -m), check_for_indels (bin/ntsynt_synteny.py:364-409; bp = --bp) and
 * filter_synteny_blocks (:411-426; min_mx = 4).  Outputs, unordered, each with room for cap >= n_runs +
 * nts_graph_big_count + 1 entries: surviving blocks (b_lo..b_hi traversed in direction b_dir, bit a of b_plus = '+'
 * in assembly a), deleted vertex intervals (r_lo..r_hi), cut pairs (c, c+1); counts[3] = how many of each. */
int nts_graph_runs_to_blocks(nts_graph* g, const int64_t* starts, const int64_t* ends, uint64_t n_runs, uint32_t bp,
                             double m_pct, uint32_t min_mx, uint32_t* b_lo, uint32_t* b_hi, uint32_t* b_plus, int8_t* b_dir,
                             uint32_t* r_lo, uint32_t* r_hi, uint32_t* cuts, uint64_t counts[3], uint64_t cap);

/* ---- native host-side pieces of the graph stage (no device work; operate on the caller's host arrays) ------------
 * nts_host_walk_paths: find_paths (subprojects/ntJoin/bin/ntjoin.py:114-151) over the components of the
 * weight-filtered graph that contain an edge other than (i, i+1).  nbr = [n_vertices x 2] neighbours or -1;
 * (starts, ends) = the n_runs maximal (i, i+1) chains over base vertices [0, V0), ascending; sv = the n_sv vertices
 * holding such an edge, ascending; opos = position of every vertex in the orienting assembly.  Paths come back as
 * segments (lo, hi, dir = +1 / -1) with path p owning segments [path_off[p], path_off[p+1]); seg_cap >= 2 * n_sv + 2. */
int nts_host_walk_paths(const int32_t* nbr, int64_t V0, const int64_t* starts, const int64_t* ends, int64_t n_runs,
                        const int64_t* sv, int64_t n_sv, const int64_t* opos, int64_t* seg_lo, int64_t* seg_hi,
                        int8_t* seg_dir, int64_t* path_off, int64_t seg_cap, int64_t* n_paths, int64_t* n_segs);
/* the same with the orienting positions given only where a path can end: opos_ids (ascending) = the sparse vertices
 * and both ends of the runs that hold them, opos_vals their positions (the lean form keeps positions on the device) */
int nts_host_walk_paths_sparse(const int32_t* nbr, int64_t V0, const int64_t* starts, const int64_t* ends, int64_t n_runs,
                               const int64_t* sv, int64_t n_sv, const int64_t* opos_ids, const int64_t* opos_vals, int64_t n_opos,
                               int64_t* seg_lo, int64_t* seg_hi, int8_t* seg_dir, int64_t* path_off, int64_t seg_cap,
                               int64_t* n_paths, int64_t* n_segs);
/* nts_host_paths_to_blocks: find_synteny_blocks + check_for_indels + filter_synteny_blocks (bin/ntsynt_synteny.py:66-106,
 * 364-426; bin/synteny_block.py:48-65) for the host-walked paths of a round, given as segments with their range sums and
 * a small table of the positions / contigs of the vertices involved; see csrc/nts_hostgraph.cu for the argument list. */
int nts_host_paths_to_blocks(int64_t n_paths, const int64_t* path_off, const int64_t* seg_lo, const int64_t* seg_hi,
                             const int8_t* seg_dir, uint32_t G, const int64_t* up, const int64_t* down, const int64_t* cuts,
                             const int64_t* cut_off, const int64_t* pos, const int32_t* ctg, int64_t bp, double m_pct,
                             int64_t min_mx, int64_t cap, int64_t* b_off, int64_t* b_n, int64_t* b_first, int64_t* b_last,
                             int8_t* b_ori, int32_t* b_ctg, int64_t* b_fpos, int64_t* b_lpos, int64_t* o_lo, int64_t* o_hi,
                             int8_t* o_dir, int64_t* r_lo, int64_t* r_hi, int64_t* e_u, int64_t* e_v, int64_t counts[4]);
/* nts_host_simplify: run_graph_simplification on the round-0 graph (bin/ntsynt_synteny.py:548-590), candidate edges
 * visited in build_graph's edge-id order (subprojects/ntJoin/bin/ntjoin_utils.py:97-115).  cand = the n_cand vertices
 * with exactly three distinct neighbours, ascending; rank / inv = [G x V]; ctg = [G x ctg_stride].  For every edge
 * that fires: bump_s < bump_t (the edge that becomes full weight) and removed (the triangle's third vertex). */
int nts_host_simplify(const int64_t* cand, int64_t n_cand, const uint32_t* rank, const uint32_t* inv, const int32_t* ctg,
                      int64_t ctg_stride, int64_t V, uint32_t G, int64_t* bump_s, int64_t* bump_t, int64_t* removed,
                      int64_t out_cap, int64_t* n_out);
/* the same from neighbourhoods extracted on the device (nts_graph_neigh): left / right / rk are [n_cand x G] */
int nts_host_simplify_neigh(const int64_t* cand, int64_t n_cand, const int64_t* left, const int64_t* right, const int64_t* rk,
                            uint32_t G, int64_t* bump_s, int64_t* bump_t, int64_t* removed, int64_t out_cap, int64_t* n_out);

/* ---- native FASTA ingest (host code; replaces btllib::SeqReader, src/ntsynt_make_common_bf.cpp:32-36,125,143, and
 * `samtools faidx`, bin/ntsynt_run_pipeline.smk:48-53) --------------------------------------------------------------
 * nts_fasta_scan: records of a FASTA held in memory.  Per record: name = first whitespace-delimited token of the
 * header (name_off / name_len into buf), n_bases, [seq_off, seq_end) = file span of its sequence lines, linebases /
 * linewidth of its first non-empty sequence line (the .fai columns), uniform = 1 iff every sequence line but the last
 * has that width.  *n_records may exceed cap: call again with larger arrays. */
int nts_fasta_scan(const char* buf, uint64_t n, uint64_t cap, uint64_t* name_off, uint32_t* name_len, uint64_t* n_bases,
                   uint64_t* seq_off, uint64_t* seq_end, uint32_t* linebases, uint32_t* linewidth, uint8_t* uniform,
                   uint64_t* n_records);
/* the same records found by n_threads threads (0 = all cores): headers by buffer slice, record bodies in ~8 MB pieces */
int nts_fasta_scan_mt(const char* buf, uint64_t n, uint64_t cap, uint64_t* name_off, uint32_t* name_len, uint64_t* n_bases,
                   uint64_t* seq_off, uint64_t* seq_end, uint32_t* linebases, uint32_t* linewidth, uint8_t* uniform,
                   uint64_t* n_records, uint32_t n_threads);
/* nts_gz_inflate: gzip bytes (one member or several, zero padding after a member allowed) -> plain bytes, whole buffer in,
 * whole buffer out; the decompressor in front of the FASTA reader (btllib::SeqReader pipes .gz input through one,
 * src/ntsynt_make_common_bf.cpp:32-36,125,143).  *n_out = bytes written.  Returns NTS_OK, 1 when `cap` is too small (call
 * again with a larger buffer; the ISIZE trailer of a single-member file is its exact size), NTS_ERR_STATE for a truncated
 * or corrupt stream.  verify_crc: check every member's CRC-32 (computed by a second thread behind the decoder). */
int nts_gz_inflate(const uint8_t* in, uint64_t n_in, uint8_t* out, uint64_t cap, uint64_t* n_out, int verify_crc);
/* the same by n_threads threads (0 = all cores): a member of >= 16 MB is cut into chunks; every chunk but the first finds a
 * block boundary by trial, is decoded without its history into 16-bit symbols (what it copies out of the unknown 32 KB in
 * front of it stays symbolic) and is resolved once its predecessor is done.  Same checks, same result. */
int nts_gz_inflate_mt(const uint8_t* in, uint64_t n_in, uint8_t* out, uint64_t cap, uint64_t* n_out, int verify_crc,
                      uint32_t n_threads);
/* members whose spans are known without decoding (BGZF, `bgzip`: every member names its compressed size): member i is
 * in[member_off[i] .. member_off[i + 1]); the members are decoded independently, a contiguous range per thread. */
int nts_gz_inflate_members(const uint8_t* in, const uint64_t* member_off, uint64_t n_members, uint8_t* out, uint64_t cap,
                           uint64_t* n_out, int verify_crc, uint32_t n_threads);
/* nts_fasta_pack: 2-bit pack every record with n_threads threads (0 = hardware concurrency; records in parallel, long
 * uniform records split at 4 Mbp).  word_off[r] = even offset of record r in words_out (zero-initialised, sum of
 * nts_packed_words(n_bases[r]) words); N runs in record coordinates, record r owning [nrun_off[r], nrun_off[r+1]).
 * If *n_nruns > nrun_cap only the count is valid: call again with larger arrays. */
int nts_fasta_pack(const char* buf, uint64_t n_records, const uint64_t* n_bases, const uint64_t* seq_off, const uint64_t* seq_end,
                   const uint32_t* linebases, const uint32_t* linewidth, const uint8_t* uniform, const uint64_t* word_off,
                   uint64_t* words_out, uint64_t* nrun_off, uint64_t* nrun_start, uint64_t* nrun_len, uint64_t nrun_cap,
                   uint64_t* n_nruns, uint32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif /* NTSYNT_B200_H */
