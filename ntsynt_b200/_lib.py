"""ctypes binding of libntsynt_b200.so (include/ntsynt_b200.h).  No CPU fallback: if the
library is missing it is built with nvcc; if it cannot be loaded the import fails loudly."""
import ctypes as C
import os

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))


class NtsError(RuntimeError):
    "A libntsynt_b200 call returned a negative status."

    def __init__(self, status, message):
        super().__init__(f"[nts status {status}] {message}")
        self.status = status


class SynthSeg(C.Structure):
    _fields_ = [("dst_contig", C.c_uint32), ("anc_contig", C.c_int32), ("dst_start", C.c_uint64),
                ("anc_start", C.c_uint64), ("len", C.c_uint64), ("strand", C.c_int32), ("pad", C.c_uint32)]


def _load():
    path = os.environ.get("NTSYNT_B200_LIB") or _build.LIB
    if not os.path.exists(path) or (not os.environ.get("NTSYNT_B200_LIB") and _build.needs_build()):
        path = _build.build_lib()
    try:
        return C.CDLL(path, mode=C.RTLD_GLOBAL)
    except OSError as exc:
        raise ImportError(f"cannot load {path}: {exc}. ntsynt_b200 has no CPU fallback; build it with "
                          f"`python -m ntsynt_b200.build` (needs nvcc).") from exc


lib = _load()

u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
vp, vpp = C.c_void_p, C.POINTER(C.c_void_p)
i64p = C.POINTER(C.c_int64)

_SIGS = {
    "nts_version": (C.c_char_p, []),
    "nts_last_error": (C.c_char_p, []),
    "nts_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "nts_ctx_create": (C.c_int, [C.c_int, vpp]),
    "nts_ctx_destroy": (None, [vp]),
    "nts_ctx_sync": (C.c_int, [vp]),
    "nts_timer_start": (C.c_int, [vp]),
    "nts_timer_stop": (C.c_int, [vp, C.POINTER(C.c_float)]),
    "nts_launch_count": (C.c_uint64, [vp]),
    "nts_sketch_escalated": (C.c_uint64, [vp]),
    "nts_sketch_queried_all": (C.c_uint64, [vp]),
    "nts_part_inserts": (C.c_uint64, [vp]),
    "nts_part_overflow_items": (C.c_uint64, [vp]),
    "nts_mem_info": (C.c_int, [vp, u64p, u64p]),
    "nts_prof_enable": (C.c_int, [vp, C.c_int]),
    "nts_prof_reset": (C.c_int, [vp]),
    "nts_prof_count": (C.c_int, []),
    "nts_prof_name": (C.c_char_p, [C.c_int]),
    "nts_prof_get": (C.c_int, [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), u64p]),
    "nts_xfer_bytes": (C.c_int, [vp, u64p, u64p]),
    "nts_host_alloc": (C.c_int, [C.c_uint64, vpp]),
    "nts_host_free": (None, [vp]),
    "nts_packed_words": (C.c_uint64, [C.c_uint64]),
    "nts_pack_ascii": (C.c_int, [C.c_char_p, C.c_uint64, u64p, u64p, u64p, C.c_uint64, u64p]),
    "nts_unpack_ascii": (C.c_int, [u64p, C.c_uint64, C.c_uint64, C.c_char_p]),
    "nts_genome_upload": (C.c_int, [vp, C.c_uint32, u64p, u64p, u64p, C.c_uint64, u64p, u64p, u64p, vpp]),
    "nts_genome_upload_async": (C.c_int, [vp, C.c_uint32, u64p, u64p, u64p, C.c_uint64, u64p, u64p, u64p, vpp]),
    "nts_genome_destroy": (None, [vp]),
    "nts_genome_size": (C.c_uint64, [vp]),
    "nts_genome_contigs": (C.c_uint32, [vp]),
    "nts_genome_download_contig": (C.c_int, [vp, C.c_uint32, u64p]),
    "nts_genome_nruns": (C.c_int, [vp, u64p, u64p, u64p, C.c_uint64, u64p]),
    "nts_genome_synthesize": (C.c_int, [vp, C.c_uint32, u64p, C.POINTER(SynthSeg), C.c_uint64, C.c_uint64, C.c_uint64,
                                        C.c_double, C.c_uint32, C.c_double, vpp]),
    "nts_synth_ancestor_base": (C.c_int, [C.c_uint64, C.c_uint32, C.c_double, C.c_uint32, C.c_uint64]),
    "nts_bf_bytes": (C.c_uint64, [C.c_int64, C.c_double]),
    "nts_bf_create": (C.c_int, [vp, C.c_uint64, vpp]),
    "nts_bf_destroy": (None, [vp]),
    "nts_bf_size_bytes": (C.c_uint64, [vp]),
    "nts_bf_clear": (C.c_int, [vp]),
    "nts_bf_insert_genome": (C.c_int, [vp, vp, C.c_uint32]),
    "nts_bf_set_genome": (C.c_int, [vp, vp, C.c_uint32]),
    "nts_bf_and": (C.c_int, [vp, vp]),
    "nts_bf_or": (C.c_int, [vp, vp]),
    "nts_bf_build_common": (C.c_int, [vp, vp, vpp, C.c_uint32, C.c_uint32]),
    "nts_bf_build_common_lazy": (C.c_int, [vp, vp, vpp, C.c_uint32, C.c_uint32, C.POINTER(C.c_int)]),
    "nts_bf_insert_repeats": (C.c_int, [vp, vp, vp, C.c_uint32]),
    "nts_bf_popcount": (C.c_int, [vp, u64p]),
    "nts_bf_download": (C.c_int, [vp, u8p]),
    "nts_bf_upload": (C.c_int, [vp, u8p]),
    "nts_sketch": (C.c_int, [vp, vp, vp, vp, C.c_uint32, C.c_uint32, u64p, u64p, u64p, vpp]),
    "nts_sketch2": (C.c_int, [vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32, u64p, u64p, u64p, vpp]),
    "nts_mxs_destroy": (None, [vp]),
    "nts_mxs_count": (C.c_uint64, [vp]),
    "nts_mxs_download": (C.c_int, [vp, u64p, u32p, u32p]),
    "nts_hash_contig": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, u64p, u8p]),
    "nts_mxs_upload": (C.c_int, [vp, C.c_uint64, u64p, u32p, u32p, vpp]),
    "nts_nccl_unique_id": (C.c_int, [u8p]),
    "nts_nccl_init": (C.c_int, [vp, u8p, C.c_int, C.c_int, vpp]),
    "nts_nccl_destroy": (None, [vp]),
    "nts_nccl_barrier": (C.c_int, [vp]),
    "nts_nccl_world": (C.c_int, [vp]),
    "nts_nccl_rank": (C.c_int, [vp]),
    "nts_bf_allreduce_and": (C.c_int, [vp, vp]),
    "nts_bf_allreduce_or": (C.c_int, [vp, vp]),
    "nts_bf_ipc_handle": (C.c_int, [vp, u8p]),
    "nts_p2p_open": (C.c_int, [vp, u8p, C.c_int, C.c_int, vpp]),
    "nts_p2p_close": (None, [vp]),
    "nts_p2p_reduce_scatter": (C.c_int, [vp, C.c_int]),
    "nts_p2p_all_gather": (C.c_int, [vp]),
    "nts_p2p_reduce_and_of_or": (C.c_int, [vpp, C.c_uint32, vp]),
    "nts_bin_prepare": (C.c_int, [vp, C.c_int, C.c_uint64]),
    "nts_bin_genome": (C.c_int, [vp, vp, C.c_uint32, C.c_int]),
    "nts_bin_overflow": (C.c_int, [vp, C.c_int, u64p]),
    "nts_bin_ipc_handles": (C.c_int, [vp, C.c_int, u8p, u8p]),
    "nts_binpeer_open": (C.c_int, [vp, u8p, u8p, C.c_int, C.c_int, C.c_int, vpp]),
    "nts_binpeer_close": (None, [vp]),
    "nts_bf_apply_owned": (C.c_int, [vp, vp, C.c_int, C.c_uint64, C.c_uint64]),
    "nts_bf_range_op": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, C.c_int]),
    "nts_p2p_slice": (C.c_int, [vp, C.c_int, u64p, u64p]),
    "nts_mxs_drop_in_bf": (C.c_int, [vp, vp, vp, vp, C.c_uint32, vpp]),
    "nts_mxs_contig_offsets": (C.c_int, [vp, C.c_uint32, u64p]),
    "nts_mxs_concat": (C.c_int, [vp, vpp, u64p, u64p, C.c_uint64, vpp]),
    "nts_mxs_allgather": (C.c_int, [vp, vp, u64p, vpp]),
    "nts_graph_build": (C.c_int, [vp, vpp, C.c_uint32, C.c_uint32, vpp]),
    "nts_graph_destroy": (None, [vp]),
    "nts_graph_vertices": (C.c_uint64, [vp]),
    "nts_graph_download_vertices": (C.c_int, [vp, u64p, u32p, u32p, u32p, u8p, u8p]),
    "nts_graph_download_links": (C.c_int, [vp, u32p, u32p, u32p, u32p]),
    "nts_graph_download_cums": (C.c_int, [vp, u32p, u32p]),
    "nts_graph_download_host_arrays": (C.c_int, [vp, C.c_uint64, u64p, C.POINTER(C.c_longlong), C.POINTER(C.c_int32),
                                                 C.POINTER(C.c_int32), u8p]),
    "nts_graph_sparse_lists": (C.c_int, [vp, C.c_uint32, u32p, u32p, u32p, u64p]),
    "nts_host_walk_paths": (C.c_int, [C.POINTER(C.c_int32), C.c_int64, i64p, i64p, C.c_int64, i64p, C.c_int64, i64p, i64p, i64p,
                                      C.POINTER(C.c_int8), i64p, C.c_int64, i64p, i64p]),
    "nts_host_walk_paths_sparse": (C.c_int, [C.POINTER(C.c_int32), C.c_int64, i64p, i64p, C.c_int64, i64p, C.c_int64, i64p, i64p,
                                             C.c_int64, i64p, i64p, C.POINTER(C.c_int8), i64p, C.c_int64, i64p, i64p]),
    "nts_host_paths_to_blocks": (C.c_int, [C.c_int64, i64p, i64p, i64p, C.POINTER(C.c_int8), C.c_uint32, i64p, i64p, i64p, i64p,
                                           i64p, C.POINTER(C.c_int32), C.c_int64, C.c_double, C.c_int64, C.c_int64,
                                           i64p, i64p, i64p, i64p, C.POINTER(C.c_int8), C.POINTER(C.c_int32), i64p, i64p, i64p, i64p,
                                           C.POINTER(C.c_int8), i64p, i64p, i64p, i64p, i64p]),
    "nts_host_simplify": (C.c_int, [i64p, C.c_int64, u32p, u32p, C.POINTER(C.c_int32), C.c_int64, C.c_int64, C.c_uint32,
                                    i64p, i64p, i64p, C.c_int64, i64p]),
    "nts_fasta_scan": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, u64p, u32p, u64p, u64p, u64p, u32p, u32p, u8p, u64p]),
    "nts_fasta_scan_mt": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, u64p, u32p, u64p, u64p, u64p, u32p, u32p, u8p, u64p, C.c_uint32]),
    "nts_gz_inflate": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, u64p, C.c_int]),
    "nts_gz_inflate_mt": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, u64p, C.c_int, C.c_uint32]),
    "nts_gz_inflate_members": (C.c_int, [C.c_void_p, u64p, C.c_uint64, C.c_void_p, C.c_uint64, u64p, C.c_int, C.c_uint32]),
    "nts_fasta_pack": (C.c_int, [C.c_void_p, C.c_uint64, u64p, u64p, u64p, u32p, u32p, u8p, u64p, u64p, u64p, u64p, u64p,
                                 C.c_uint64, u64p, C.c_uint32]),
    "nts_graph_lookup": (C.c_int, [vp, u64p, C.c_uint64, u32p]),
    "nts_graph_edges": (C.c_int, [vp, u64p]),
    "nts_graph_download_edges": (C.c_int, [vp, u32p, u32p, u32p]),
    "nts_graph_gather": (C.c_int, [vp, C.c_int, i64p, C.c_uint64, vp]),
    "nts_graph_range_sums": (C.c_int, [vp, i64p, i64p, C.c_uint64, i64p, i64p]),
    "nts_graph_neigh": (C.c_int, [vp, i64p, C.c_uint64, i64p, i64p, i64p]),
    "nts_graph_download_links_nbr": (C.c_int, [vp, C.c_uint64, C.POINTER(C.c_int32), u8p]),
    "nts_graph_set_links": (C.c_int, [vp, i64p, u8p, C.c_uint64]),
    "nts_graph_runs": (C.c_int, [vp, i64p, i64p, u64p]),
    "nts_graph_refine_filter": (C.c_int, [vp, vpp, u32p, u32p, C.c_uint32, u32p, C.c_uint32, u64p, u32p, C.c_uint32, i64p, i64p, u64p,
                                          u64p, u64p, u64p, u32p, u32p, u32p, C.c_uint64]),
    "nts_graph_big_count": (C.c_int, [vp, C.c_uint32, u64p]),
    "nts_graph_runs_to_blocks": (C.c_int, [vp, i64p, i64p, C.c_uint64, C.c_uint32, C.c_double, C.c_uint32, u32p, u32p, u32p,
                                           C.POINTER(C.c_int8), u32p, u32p, u32p, u64p, C.c_uint64]),
    "nts_host_simplify_neigh": (C.c_int, [i64p, C.c_int64, i64p, i64p, i64p, C.c_uint32, i64p, i64p, i64p, C.c_int64, i64p]),
}


def declared_symbols():
    "every symbol include/ntsynt_b200.h declares (parsed from the header itself)"
    import re
    hdr = os.path.join(_HERE, "..", "include", "ntsynt_b200.h")
    with open(hdr, encoding="utf-8") as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nts_[a-z0-9_]+)\s*\(", text)))


for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


def check(status):
    if status != 0:
        raise NtsError(status, lib.nts_last_error().decode(errors="replace"))


def ptr(arr, ctype):
    "pointer to a contiguous numpy array (or None)"
    if arr is None:
        return None
    return arr.ctypes.data_as(C.POINTER(ctype))
