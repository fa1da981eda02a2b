"""ctypes binding of the C ABI declared in include/percept_smoother_c.h.

`Api(lib, prefix)` binds every `<prefix>*` entry point of a shared library that
implements that header; `Ctx` wraps one `pms_ctx*`.  The product binds
`percept_b200/lib/libpms.so` with prefix ``pms_`` (see percept_b200/__init__.py).
tests/ bind the CPU checker's twin ABI (a different symbol prefix) through the same class so
that parity tests issue the identical call sequence to both.
"""
import ctypes as C
import numpy as np

AOS, SOA = 0, 1
TET4, HEX8 = 4, 8

FIELDS3 = ("coordinates", "coordinates_N", "coordinates_NM1", "coordinates_lagged",
           "cg_g", "cg_r", "cg_d", "cg_s")
FIELDS1 = ("cg_edge_length", "num_adj_elems")


class State(C.Structure):
    """pms_state: mirror of ReferenceMeshSmootherBase.hpp:71-101 state scalars."""
    _fields_ = [(n, C.c_double) for n in
                ("dmax", "dmax_relative", "dnew", "dold", "d0", "dmid", "alpha", "alpha_0",
                 "grad_norm_scaled", "total_metric", "global_metric")] + \
               [(n, C.c_int) for n in ("stage", "iter", "untangled", "converged")] + \
               [(n, C.c_uint64) for n in ("num_invalid", "num_nodes", "n_metric_evals",
                                          "n_gradient_evals", "n_line_search_halvings", "restarted")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class SmootherError(RuntimeError):
    """Non-zero status from the C ABI (the reference throws std::runtime_error)."""

    def __init__(self, code, msg):
        super().__init__(f"[status {code}] {msg}")
        self.code = code


_I3 = C.c_int * 3
_U3 = C.c_uint * 3
_dp = C.POINTER(C.c_double)


def _dptr(a):
    return a.ctypes.data_as(_dp)


class Api:
    # name -> argtypes (after the ctx pointer); all return int unless noted
    _SIGS = {
        "set_option": [C.c_char_p, C.c_double],
        "add_structured_block": [_U3, C.POINTER(C.c_int)],
        "block_add_boundary_condition": [C.c_int, _I3, _I3],
        "block_add_zone_connectivity": [C.c_int, C.c_int, _I3, _I3, _I3, _I3, _I3],
        "set_unstructured": [C.c_int, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p],
        "set_fixed_mask": [C.c_void_p],
        "select_boundary_sgrid": [],
        "commit": [],
        "num_nodes_local": [C.POINTER(C.c_size_t)],
        "num_elements_local": [C.POINTER(C.c_size_t)],
        "get_node_elem_csr": [C.c_void_p, C.c_void_p],
        "field_upload": [C.c_char_p, C.c_int, _dp, C.c_int],
        "field_download": [C.c_char_p, C.c_int, _dp, C.c_int],
        "copy_field": [C.c_char_p, C.c_char_p],
        "nodal_field_axpby": [C.c_double, C.c_char_p, C.c_double, C.c_char_p],
        "nodal_field_axpbypgz": [C.c_double, C.c_char_p, C.c_double, C.c_char_p, C.c_double, C.c_char_p],
        "nodal_field_dot": [C.c_char_p, C.c_char_p, _dp],
        "nodal_field_set_value": [C.c_char_p, C.c_double],
        "get_number_nodes": [C.POINTER(C.c_size_t)],
        "get_edge_lengths": [],
        "total_metric": [C.c_int, C.c_double, _dp, C.POINTER(C.c_size_t)],
        "get_gradient": [C.c_int],
        "metric_and_gradient": [C.c_int, _dp, C.POINTER(C.c_size_t)],
        "count_invalid_elements": [C.POINTER(C.c_size_t)],
        "get_alpha_0": [_dp],
        "update_node_positions": [C.c_double, _dp, _dp],
        "get_element_metrics": [C.c_void_p, C.c_void_p],
        "state_init": [C.POINTER(State)],
        "run_one_iteration": [C.POINTER(State), C.c_double, _dp],
        "line_search": [C.POINTER(State), C.c_double, C.POINTER(C.c_int), _dp],
        "run_algorithm": [C.c_int, C.c_double, C.c_char_p, C.POINTER(State)],
        "add_analytic_surface": [C.c_int, _dp, C.POINTER(C.c_int)],
        "set_node_classification": [C.c_void_p, C.c_void_p],
        "snap_nodes": [_dp],
        "get_surface_normals": [_dp],
        "refine_structured_topology": [C.c_void_p, C.POINTER(C.c_ulonglong)],
        "refine_structured_field": [C.c_char_p, C.c_void_p, C.c_char_p],
    }
    # product-only entry points (absent from the oracle twin)
    _SIGS_PRODUCT = {
        "field_download_async": [C.c_char_p, C.c_int, _dp],
        "field_upload_async": [C.c_char_p, C.c_int, _dp],
        "field_upload_finish": [],
        "transfer_wait": [],
        "total_metric_multi": [C.c_int, C.c_int, _dp, _dp, C.POINTER(C.c_size_t)],
        "field_device_ptr": [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)],
        "field_release_ptr": [C.c_char_p],
        "snap_point": [C.c_size_t, _dp, C.c_int],
        "comm_init": [C.c_void_p, C.c_int, C.c_int],
        "set_halo": [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
        "halo_exchange": [C.c_char_p],
        "set_interface": [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
        "launch_count": [C.POINTER(C.c_uint64)],
        "stream": [C.POINTER(C.c_void_p)],
        "synchronize": [],
        "kernel_time_ms": [C.c_char_p, _dp, C.POINTER(C.c_uint64)],
        "kernel_time_reset": [],
        "set_timing": [C.c_int],
    }

    def __init__(self, lib, prefix, product):
        self.lib, self.prefix, self.product = lib, prefix, product
        self.fn = {}
        sigs = dict(self._SIGS)
        if product:
            sigs.update(self._SIGS_PRODUCT)
        for name, args in sigs.items():
            f = getattr(lib, prefix + name)  # AttributeError if the library lacks a declared symbol
            f.argtypes = [C.c_void_p] + args
            f.restype = C.c_int
            self.fn[name] = f
        self._create = getattr(lib, prefix + "create")
        self._create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        self._create.restype = C.c_int
        self._destroy = getattr(lib, prefix + "destroy")
        self._destroy.argtypes = [C.c_void_p]
        self._destroy.restype = None
        self._last_error = getattr(lib, prefix + "last_error")
        self._last_error.argtypes = [C.c_void_p]
        self._last_error.restype = C.c_char_p

    def create(self, device=0):
        return Ctx(self, device)


class Ctx:
    """One smoother context (one GPU / one oracle instance)."""

    def __init__(self, api, device=0):
        self.api = api
        p = C.c_void_p()
        rc = api._create(C.byref(p), int(device))
        if rc != 0 or not p:
            msg = api._last_error(p).decode() if p else "create failed (no CUDA device?)"
            if p:
                api._destroy(p)
            raise SmootherError(rc, msg)
        self.p = p
        self.block_sizes = []

    def close(self):
        if getattr(self, "p", None):
            self.api._destroy(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, name, *args):
        rc = self.api.fn[name](self.p, *args)
        if rc != 0:
            raise SmootherError(rc, self.api._last_error(self.p).decode())

    # ---- mesh ----
    def set_option(self, key, value):
        self._call("set_option", key.encode(), float(value))

    def add_structured_block(self, node_size):
        bid = C.c_int(-1)
        self._call("add_structured_block", _U3(*[int(v) for v in node_size]), C.byref(bid))
        self.block_sizes.append(tuple(int(v) for v in node_size))
        return bid.value

    def add_boundary_condition(self, block, beg, end):
        self._call("block_add_boundary_condition", int(block), _I3(*beg), _I3(*end))

    def add_fixture_1_bcs(self, block):
        """The 6 boundary conditions of StructuredBlock::fixture_1 (StructuredBlock.cpp:158-163)."""
        n = self.block_sizes[block]
        for beg, end in (((1, 1, 1), (1, n[1], n[2])), ((n[0], 1, 1), (n[0], n[1], n[2])),
                         ((1, 1, 1), (n[0], 1, n[2])), ((1, n[1], 1), (n[0], n[1], n[2])),
                         ((1, 1, 1), (n[0], n[1], 1)), ((1, 1, n[2]), (n[0], n[1], n[2]))):
            self.add_boundary_condition(block, beg, end)

    def add_zone_connectivity(self, block, donor, transform, owner_beg, owner_end, donor_beg, donor_end):
        self._call("block_add_zone_connectivity", int(block), int(donor), _I3(*transform), _I3(*owner_beg),
                   _I3(*owner_end), _I3(*donor_beg), _I3(*donor_end))

    def set_unstructured(self, topology, nnodes, conn, elem_owned=None, node_owned=None):
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        nelems = conn.size // topology
        eo = None if elem_owned is None else np.ascontiguousarray(elem_owned, dtype=np.uint8)
        no = None if node_owned is None else np.ascontiguousarray(node_owned, dtype=np.uint8)
        self._call("set_unstructured", int(topology), int(nnodes), int(nelems), conn.ctypes.data,
                   None if eo is None else eo.ctypes.data, None if no is None else no.ctypes.data)

    def set_fixed_mask(self, fixed):
        if fixed is None:
            self._call("set_fixed_mask", None)
        else:
            f = np.ascontiguousarray(fixed, dtype=np.uint8)
            self._call("set_fixed_mask", f.ctypes.data)

    def select_boundary_sgrid(self):
        self._call("select_boundary_sgrid")

    def commit(self):
        self._call("commit")

    # ---- surface smoothing hooks (MeshGeometry stand-ins) ----
    def add_analytic_surface(self, kind, params):
        """kind: 0 plane (point, normal), 1 sphere (centre, radius), 2 cylinder (axis point, axis dir, radius)."""
        p = np.zeros(7)
        p[:len(params)] = params
        sid = C.c_int(-1)
        self._call("add_analytic_surface", int(kind), _dptr(p), C.byref(sid))
        return sid.value

    def set_node_classification(self, dof, surface_id):
        d = np.ascontiguousarray(dof, dtype=np.int8)
        s = np.ascontiguousarray(surface_id, dtype=np.int32)
        self._call("set_node_classification", d.ctypes.data, s.ctypes.data)

    def snap_nodes(self):
        d = C.c_double(0.0)
        self._call("snap_nodes", C.byref(d))
        return d.value

    def get_surface_normals(self):
        out = np.zeros((3, self.num_nodes_local))
        self._call("get_surface_normals", _dptr(out))
        return out

    # ---- StructuredGridRefiner (structured/StructuredGridRefiner.cpp:40-170) ----
    def refine_topology(self, fine):
        """Fill the empty ctx `fine` with the 2x refinement of this grid's blocks; returns the refined cell count."""
        n = C.c_ulonglong(0)
        rc = self.api.fn["refine_structured_topology"](self.p, fine.p, C.byref(n))
        if rc != 0:
            raise SmootherError(rc, self.api._last_error(fine.p).decode())
        fine.block_sizes = [tuple(2 * v - 1 for v in b) for b in self.block_sizes]
        return int(n.value)

    def refine_field(self, src_field, fine, dst_field):
        """2x prolongation of a 3-component nodal field onto the committed refined grid `fine`."""
        rc = self.api.fn["refine_structured_field"](self.p, src_field.encode(), fine.p, dst_field.encode())
        if rc != 0:
            raise SmootherError(rc, self.api._last_error(fine.p).decode())

    @property
    def num_nodes_local(self):
        n = C.c_size_t()
        self._call("num_nodes_local", C.byref(n))
        return n.value

    @property
    def num_elements_local(self):
        n = C.c_size_t()
        self._call("num_elements_local", C.byref(n))
        return n.value

    def node_elem_csr(self, npe):
        nn, ne = self.num_nodes_local, self.num_elements_local
        rp = np.zeros(nn + 1, dtype=np.int64)
        en = np.zeros(ne * npe, dtype=np.int64)
        self._call("get_node_elem_csr", rp.ctypes.data, en.ctypes.data)
        return rp, en

    # ---- fields ----
    def _nblock(self, block):
        if block < 0:
            return self.num_nodes_local
        n = self.block_sizes[block]
        return n[0] * n[1] * n[2]

    def upload(self, name, arr, block=-1, layout=SOA):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        nc = 3 if name in FIELDS3 else 1
        assert a.size == nc * self._nblock(block), (name, a.shape, self._nblock(block))
        self._call("field_upload", name.encode(), int(block), _dptr(a), int(layout))

    def download(self, name, block=-1, layout=SOA, out=None):
        nc = 3 if name in FIELDS3 else 1
        n = self._nblock(block)
        if out is None:
            out = np.empty((nc, n) if layout == SOA else (n, nc), dtype=np.float64)
        self._call("field_download", name.encode(), int(block), _dptr(out), int(layout))
        return out

    def download_async(self, name, out, block=-1):
        """SoA download into the (pinned) array `out` on a second stream; call transfer_wait() before reading it."""
        nc = 3 if name in FIELDS3 else 1
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size == nc * self._nblock(block)
        self._call("field_download_async", name.encode(), int(block), _dptr(out))

    def upload_async(self, name, arr, block=-1):
        """start host -> device staging on a side stream; the field changes at upload_finish()."""
        nc = 3 if name in FIELDS3 else 1
        assert arr.dtype == np.float64 and arr.flags.c_contiguous and arr.size == nc * self._nblock(block)
        self._call("field_upload_async", name.encode(), int(block), _dptr(arr))

    def upload_finish(self):
        self._call("field_upload_finish")

    def transfer_wait(self):
        self._call("transfer_wait")

    def copy_field(self, dst, src):
        self._call("copy_field", dst.encode(), src.encode())

    def axpby(self, alpha, x, beta, y):
        self._call("nodal_field_axpby", float(alpha), x.encode(), float(beta), y.encode())

    def axpbypgz(self, alpha, x, beta, y, gamma, z):
        self._call("nodal_field_axpbypgz", float(alpha), x.encode(), float(beta), y.encode(), float(gamma), z.encode())

    def dot(self, x, y):
        r = C.c_double()
        self._call("nodal_field_dot", x.encode(), y.encode(), C.byref(r))
        return r.value

    def set_value(self, name, v):
        self._call("nodal_field_set_value", name.encode(), float(v))

    def get_number_nodes(self):
        n = C.c_size_t()
        self._call("get_number_nodes", C.byref(n))
        return n.value

    # ---- smoother kernels ----
    def get_edge_lengths(self):
        self._call("get_edge_lengths")

    def total_metric(self, stage, alpha=0.0):
        m, n = C.c_double(), C.c_size_t()
        self._call("total_metric", int(stage), float(alpha), C.byref(m), C.byref(n))
        return m.value, n.value

    def total_metric_multi(self, stage, alphas):
        """metric and invalid-element count at every alpha of `alphas` (product only)."""
        a = np.ascontiguousarray(alphas, dtype=np.float64)
        m = np.zeros(a.size)
        n = (C.c_size_t * a.size)()
        self._call("total_metric_multi", int(stage), int(a.size), _dptr(a), _dptr(m), n)
        return m, np.array(list(n), dtype=np.int64)

    def get_gradient(self, stage):
        self._call("get_gradient", int(stage))

    def metric_and_gradient(self, stage):
        m, n = C.c_double(), C.c_size_t()
        self._call("metric_and_gradient", int(stage), C.byref(m), C.byref(n))
        return m.value, n.value

    def count_invalid_elements(self):
        n = C.c_size_t()
        self._call("count_invalid_elements", C.byref(n))
        return n.value

    def get_alpha_0(self):
        a = C.c_double()
        self._call("get_alpha_0", C.byref(a))
        return a.value

    def update_node_positions(self, alpha):
        a, b = C.c_double(), C.c_double()
        self._call("update_node_positions", float(alpha), C.byref(a), C.byref(b))
        return a.value, b.value

    def element_metrics(self):
        ne = self.num_elements_local
        m = np.zeros(ne, dtype=np.float64)
        f = np.zeros(ne, dtype=np.int32)
        self._call("get_element_metrics", m.ctypes.data, f.ctypes.data)
        return m, f

    # ---- driver ----
    def state_init(self):
        st = State()
        self._call("state_init", C.byref(st))
        return st

    def run_one_iteration(self, st, grad_norm):
        m = C.c_double()
        self._call("run_one_iteration", C.byref(st), float(grad_norm), C.byref(m))
        return m.value

    def line_search(self, st, mfac_mult=1.0):
        """line_search(restarted, mfac_mult) on its own -> (alpha, restarted); the nodes stay where they are."""
        r, a = C.c_int(0), C.c_double()
        self._call("line_search", C.byref(st), float(mfac_mult), C.byref(r), C.byref(a))
        return a.value, bool(r.value)

    def snap_point(self, node, xyz, reset=False):
        """snap_to(node, coordinate, reset): returns the snapped position."""
        v = (C.c_double * 3)(*xyz)
        self._call("snap_point", int(node), v, 1 if reset else 0)
        return [v[0], v[1], v[2]]

    def run_algorithm(self, inner_iterations, grad_norm, log_path=None):
        st = State()
        self._call("run_algorithm", int(inner_iterations), float(grad_norm),
                   None if log_path is None else str(log_path).encode(), C.byref(st))
        return st

    # ---- product-only ----
    def field_device_ptr(self, name):
        p, s = C.c_void_p(), C.c_size_t()
        self._call("field_device_ptr", name.encode(), C.byref(p), C.byref(s))
        return p.value, s.value

    def field_release_ptr(self, name):
        self._call("field_release_ptr", name.encode())

    def launch_count(self):
        n = C.c_uint64()
        self._call("launch_count", C.byref(n))
        return n.value

    def stream(self):
        p = C.c_void_p()
        self._call("stream", C.byref(p))
        return p.value or 0

    def synchronize(self):
        self._call("synchronize")

    def set_timing(self, on):
        self._call("set_timing", int(bool(on)))

    def kernel_time_reset(self):
        self._call("kernel_time_reset")

    def kernel_time_ms(self, family):
        ms, n = C.c_double(), C.c_uint64()
        self._call("kernel_time_ms", family.encode(), C.byref(ms), C.byref(n))
        return ms.value, n.value

    def comm_init(self, uid_bytes, rank, nranks):
        buf = C.create_string_buffer(bytes(uid_bytes), 128)
        self._call("comm_init", C.cast(buf, C.c_void_p), int(rank), int(nranks))

    def set_halo(self, neighbor_ranks, send_off, send_nodes, recv_off, recv_nodes):
        nr = np.ascontiguousarray(neighbor_ranks, dtype=np.int32)
        so = np.ascontiguousarray(send_off, dtype=np.int64)
        sn = np.ascontiguousarray(send_nodes, dtype=np.int32)
        ro = np.ascontiguousarray(recv_off, dtype=np.int64)
        rn = np.ascontiguousarray(recv_nodes, dtype=np.int32)
        self._call("set_halo", int(nr.size), nr.ctypes.data, so.ctypes.data, sn.ctypes.data, ro.ctypes.data,
                   rn.ctypes.data)

    def set_interface(self, neighbor_ranks, off, nodes, owned=None):
        nr = np.ascontiguousarray(neighbor_ranks, dtype=np.int32)
        of = np.ascontiguousarray(off, dtype=np.int64)
        nd = np.ascontiguousarray(nodes, dtype=np.int32)
        ow = None if owned is None else np.ascontiguousarray(owned, dtype=np.uint8)
        self._call("set_interface", int(nr.size), nr.ctypes.data, of.ctypes.data, nd.ctypes.data,
                   None if ow is None else ow.ctypes.data)

    def halo_exchange(self, name):
        self._call("halo_exchange", name.encode())
