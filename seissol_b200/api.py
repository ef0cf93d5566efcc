"""ctypes binding of the C ABI (include/seissol_b200.h -> seissol_b200/lib/libseissol_b200.so).

PyTorch is used only as the owner of device memory and streams; every compute call
goes through the C ABI into the hand-written sm_100a kernels.  There is no CPU
fallback: a missing library or a missing CUDA device raises.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# SSB_LIB selects a differently-configured build of the same library (A/B measurements, seissol_b200/build.py)
LIB_PATH = os.environ.get("SSB_LIB") or os.path.join(HERE, "lib", "libseissol_b200.so")

SYMBOLS = [
    "ssb_create", "ssb_destroy", "ssb_set_layout", "ssb_layout_reference", "ssb_query",
    "ssb_ader", "ssb_local", "ssb_ader_local", "ssb_taylor", "ssb_lts_accumulate", "ssb_neigh",
    "ssb_local_faceproj", "ssb_ader_local_faceproj", "ssb_neigh_faceproj", "ssb_taylor_faceproj", "ssb_accumulate_rows",
    "ssb_set_dr_layout", "ssb_dr_interp", "ssb_friction_lsw", "ssb_friction_rsfvw", "ssb_neigh_dr",
    "ssb_visco_ader", "ssb_visco_local", "ssb_visco_neigh", "ssb_visco_neigh_dr",
    "ssb_setup_cell_matrices", "ssb_setup_dr_matrices", "ssb_local_nodal_bc", "ssb_local_nodal_flux", "ssb_volume_energies", "ssb_fault_energies",
    "ssb_error_string", "ssb_version",
]


class SsbArray(ctypes.Structure):
    _fields_ = [("ptrs", ctypes.c_void_p), ("base", ctypes.c_void_p), ("stride_bytes", ctypes.c_longlong)]


FRICTION_DATA_FIELDS = [
    "q_plus", "q_minus", "imposed_plus", "imposed_minus", "imp_and_eta", "initial_stress",
    "mu", "slip1", "slip2", "slip_rate1", "slip_rate2", "slip_rate_magnitude", "accumulated_slip_magnitude",
    "traction1", "traction2", "rupture_time", "peak_slip_rate", "dyn_stress_time",
    "rupture_time_pending", "dyn_stress_time_pending",
    "d_c", "mu_s", "mu_d", "cohesion", "forced_rupture_time",
    "rs_a", "rs_sl0", "rs_f0", "rs_mu_w", "rs_b", "rs_sr_w", "state_variable", "convergence_inner",
    "convergence_outer", "resample_t", "initial_pressure", "nucleation_stress", "nucleation_pressure",
    "energy_slip", "energy_accumulated_slip", "frictional_energy", "space_weights", "doubled_surface_area"]


class SsbFrictionData(ctypes.Structure):
    """ssb_friction_data_t: FrictionLawData as device pointers."""
    _fields_ = [(name, ctypes.c_void_p) for name in FRICTION_DATA_FIELDS]


class SsbFrictionParams(ctypes.Structure):
    """ssb_friction_params_t."""
    _fields_ = [("delta_t", ctypes.c_double * 8), ("time_weights", ctypes.c_double * 8),
                ("full_update_time", ctypes.c_double), ("eta_p_damp", ctypes.c_double),
                ("healing_threshold", ctypes.c_double), ("t0", ctypes.c_double), ("rs_sr0", ctypes.c_double),
                ("rs_slip_rate_tolerance", ctypes.c_double), ("rs_state_tolerance", ctypes.c_double),
                ("rs_max_number_slip_rate_updates", ctypes.c_int), ("rs_number_state_variable_updates", ctypes.c_int),
                ("nucleation_count", ctypes.c_int), ("energies_from_across_fault_velocities", ctypes.c_int),
                ("nucleation_t0", ctypes.c_double * 4), ("nucleation_s0", ctypes.c_double * 4),
                ("rs_float_resolution_exit", ctypes.c_int), ("reserved", ctypes.c_int)]


class SsbError(RuntimeError):
    pass


_lib = None


def load_library():
    """Load libseissol_b200.so; fails loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SsbError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.ssb_error_string.restype = ctypes.c_char_p
    lib.ssb_version.restype = ctypes.c_char_p
    A, P, LL, I = SsbArray, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
    lib.ssb_create.argtypes = [ctypes.POINTER(P), I, I, I, I]
    lib.ssb_destroy.argtypes = [P]
    lib.ssb_set_layout.argtypes = [P, I, P, P]
    lib.ssb_layout_reference.argtypes = [P]
    lib.ssb_query.argtypes = [P, P]
    lib.ssb_ader.argtypes = [P, A, A, A, A, P, LL, P, P]
    lib.ssb_local.argtypes = [P, A, A, A, P, LL, P, P]
    lib.ssb_ader_local.argtypes = [P, A, A, A, A, P, P, LL, P, P]
    lib.ssb_taylor.argtypes = [P, A, A, P, LL, P]
    lib.ssb_neigh.argtypes = [P, A, P, P, A, A, P, P, LL, P, P]
    lib.ssb_lts_accumulate.argtypes = [P, A, A, I, LL, P]
    lib.ssb_local_faceproj.argtypes = [P, A, A, A, A, P, LL, P, P]
    lib.ssb_ader_local_faceproj.argtypes = [P, A, A, A, A, A, P, P, LL, P, P]
    lib.ssb_neigh_faceproj.argtypes = [P, A, P, P, A, A, P, P, LL, P, P]
    lib.ssb_taylor_faceproj.argtypes = [P, A, A, P, P, LL, P]
    lib.ssb_accumulate_rows.argtypes = [P, A, A, I, LL, I, P]
    lib.ssb_set_dr_layout.argtypes = [P, I]
    lib.ssb_dr_interp.argtypes = [P, A, A, P, P, P, P, P, LL, P, P]
    lib.ssb_friction_lsw.argtypes = [P, ctypes.POINTER(SsbFrictionData), ctypes.POINTER(SsbFrictionParams), LL, P]
    lib.ssb_friction_rsfvw.argtypes = [P, ctypes.POINTER(SsbFrictionData), ctypes.POINTER(SsbFrictionParams), LL, P]
    lib.ssb_neigh_dr.argtypes = [P, A, P, P, P, P, LL, P]
    lib.ssb_visco_ader.argtypes = [P, A, A, A, A, A, P, LL, P]
    lib.ssb_visco_local.argtypes = [P, A, A, A, A, A, P, LL, P]
    lib.ssb_visco_neigh.argtypes = [P, A, A, P, P, A, A, P, P, LL, P, P]
    lib.ssb_visco_neigh_dr.argtypes = [P, A, A, A, P, P, P, P, LL, P]
    lib.ssb_local_nodal_bc.argtypes = [P, A, A, A, P, P, P, P, P, P, P, P, LL, P]
    lib.ssb_local_nodal_flux.argtypes = [P, A, A, P, P, P, LL, P]
    lib.ssb_volume_energies.argtypes = [P, A, P, P, LL, P, P]
    lib.ssb_fault_energies.argtypes = [P, P, P, P, P, P, P, LL, P, P]
    lib.ssb_setup_cell_matrices.argtypes = [P, P, P, P, P, A, A, LL, P]
    lib.ssb_setup_dr_matrices.argtypes = [P, P, P, P, P, P, P, P, P, P, P, LL, P]
    _lib = lib
    return lib


def _check(code):
    if code != 0:
        raise SsbError(f"seissol_b200 error {code}: {load_library().ssb_error_string(code).decode()}")


def _check_device(t):
    """The entry points launch on the CURRENT device (include/seissol_b200.h): the tensors must live there."""
    import torch
    if t.device.index != torch.cuda.current_device():
        raise SsbError(f"tensor on {t.device} but the current CUDA device is {torch.cuda.current_device()}: "
                       "call torch.cuda.set_device / use a torch.cuda.device guard around the seissol_b200 calls")


def strided(t):
    """torch tensor [n, ...] (contiguous, on the GPU) -> base + stride descriptor."""
    assert t.is_cuda and t.is_contiguous()
    _check_device(t)
    stride = t.stride(0) * t.element_size() if t.dim() > 0 and t.shape[0] > 0 else 0
    return SsbArray(None, t.data_ptr(), stride)


def pointer_table(t):
    """torch int64 tensor of device pointers (the reference's `real**` tables)."""
    assert t.is_cuda and t.dtype.is_floating_point is False
    _check_device(t)
    return SsbArray(t.data_ptr(), None, 0)


NONE = SsbArray(None, None, 0)


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(stream):
    import torch
    s = torch.cuda.current_stream() if stream is None else stream
    return ctypes.c_void_p(s.cuda_stream)


class Context:
    """One (order, precision, equation) configuration; mirrors the reference's
    compile-time configuration (src/Config.h.in)."""

    def __init__(self, order, real_size, equation=0, mechanisms=0, reference_layout=False):
        self.lib = load_library()
        self.handle = ctypes.c_void_p()
        _check(self.lib.ssb_create(ctypes.byref(self.handle), order, real_size, equation, mechanisms))
        self.order, self.real_size = order, real_size
        if reference_layout:
            _check(self.lib.ssb_layout_reference(self.handle))
        self._query()

    def _query(self):
        out = (ctypes.c_int * 24)()
        _check(self.lib.ssb_query(self.handle, out))
        self.B, self.ld, self.qsize, self.dersize = out[0], out[1], out[2], out[3]
        self.der_offset = [out[4 + d] for d in range(self.order)]
        self.der_ld = [out[12 + d] for d in range(self.order)]
        self.e_local, self.e_neigh = out[20], out[21]
        self.P, self.ppad = out[22], out[23]

    def __del__(self):
        try:
            if self.handle:
                self.lib.ssb_destroy(self.handle)
        except Exception:
            pass

    @staticmethod
    def _coeffs(c):
        a = np.ascontiguousarray(c, dtype=np.float64)
        return a, a.ctypes.data_as(ctypes.c_void_p)

    def ader(self, dofs, idofs, derivs, local_int, coeffs, n, perm=None, stream=None):
        keep, c = self._coeffs(coeffs)
        _check(self.lib.ssb_ader(self.handle, dofs, idofs, derivs if derivs is not None else NONE,
                                 local_int, c, n, _ptr(perm), _stream(stream)))

    def local(self, dofs, idofs, local_int, face_types, n, perm=None, stream=None):
        _check(self.lib.ssb_local(self.handle, dofs, idofs, local_int, _ptr(face_types), n, _ptr(perm),
                                  _stream(stream)))

    def ader_local(self, dofs, idofs, derivs, local_int, face_types, coeffs, n, perm=None, stream=None):
        keep, c = self._coeffs(coeffs)
        _check(self.lib.ssb_ader_local(self.handle, dofs, idofs, derivs if derivs is not None else NONE,
                                       local_int, _ptr(face_types), c, n, _ptr(perm), _stream(stream)))

    def taylor(self, derivs, idofs, coeffs, n, stream=None):
        keep, c = self._coeffs(coeffs)
        _check(self.lib.ssb_taylor(self.handle, derivs, idofs, c, n, _stream(stream)))

    def lts_accumulate(self, buffers, idofs, reset, n, stream=None):
        _check(self.lib.ssb_lts_accumulate(self.handle, buffers, idofs, 1 if reset else 0, n, _stream(stream)))

    def neigh(self, dofs, idofs, neigh_int, face_types, face_rel, n, nb_idx=None, nb_ptrs=None,
              perm=None, stream=None):
        _check(self.lib.ssb_neigh(self.handle, dofs, _ptr(nb_ptrs), _ptr(nb_idx), idofs, neigh_int,
                                  _ptr(face_types), _ptr(face_rel), n, _ptr(perm), _stream(stream)))

    # -- face-projection path (include/seissol_b200.h) -------------------------------------------
    def face_proj_reals(self):
        """reals per element of the face-projection array: [4][9][B2P]."""
        b2 = self.order * (self.order + 1) // 2
        return 4 * 9 * ((b2 + 3) // 4 * 4)

    def local_faceproj(self, dofs, idofs, local_int, face_proj, face_types, n, perm=None, stream=None):
        _check(self.lib.ssb_local_faceproj(self.handle, dofs, idofs, local_int, face_proj, _ptr(face_types), n,
                                           _ptr(perm), _stream(stream)))

    def ader_local_faceproj(self, dofs, idofs, derivs, local_int, face_proj, face_types, coeffs, n, perm=None,
                            stream=None):
        keep, c = self._coeffs(coeffs)
        _check(self.lib.ssb_ader_local_faceproj(self.handle, dofs, idofs, derivs if derivs is not None else NONE,
                                                local_int, face_proj, _ptr(face_types), c, n, _ptr(perm),
                                                _stream(stream)))

    def neigh_faceproj(self, dofs, face_proj, neigh_int, face_types, face_rel, n, nb_idx=None, nb_ptrs=None,
                       perm=None, stream=None):
        _check(self.lib.ssb_neigh_faceproj(self.handle, dofs, _ptr(nb_ptrs), _ptr(nb_idx),
                                           face_proj if face_proj is not None else NONE, neigh_int,
                                           _ptr(face_types), _ptr(face_rel), n, _ptr(perm), _stream(stream)))

    def taylor_faceproj(self, derivs, proj_out, face_j, coeffs, n, stream=None):
        keep, c = self._coeffs(coeffs)
        _check(self.lib.ssb_taylor_faceproj(self.handle, derivs, proj_out, _ptr(face_j), c, n, _stream(stream)))

    def accumulate_rows(self, dst, src, reset, n, reals_per_row, stream=None):
        _check(self.lib.ssb_accumulate_rows(self.handle, dst, src, 1 if reset else 0, n, reals_per_row, _stream(stream)))

    # -- viscoelastic2 --------------------------------------------------------------------------
    def visco_ader(self, dofs, dofs_ane, idofs, idofs_ane, local_int, coeffs, n, stream=None):
        keep, c = self._coeffs(coeffs)
        _check(self.lib.ssb_visco_ader(self.handle, dofs, dofs_ane, idofs, idofs_ane, local_int, c, n, _stream(stream)))

    def visco_local(self, dofs, dofs_ane, idofs, idofs_ane, local_int, face_types, n, stream=None):
        _check(self.lib.ssb_visco_local(self.handle, dofs, dofs_ane, idofs, idofs_ane, local_int, _ptr(face_types), n,
                                        _stream(stream)))

    def visco_neigh(self, dofs, dofs_ane, idofs, neigh_int, face_types, face_rel, n, nb_idx=None, nb_ptrs=None,
                    perm=None, stream=None):
        _check(self.lib.ssb_visco_neigh(self.handle, dofs, dofs_ane, _ptr(nb_ptrs), _ptr(nb_idx), idofs, neigh_int,
                                        _ptr(face_types), _ptr(face_rel), n, _ptr(perm), _stream(stream)))

    def visco_neigh_dr(self, dofs, dofs_ane, neigh_int, godunov_ptrs, flux_solver_ptrs, side_rel, face_types, n, stream=None):
        _check(self.lib.ssb_visco_neigh_dr(self.handle, dofs, dofs_ane, neigh_int, _ptr(godunov_ptrs), _ptr(flux_solver_ptrs),
                                           _ptr(side_rel), _ptr(face_types), n, _stream(stream)))

    def local_nodal_bc(self, dofs, idofs, neigh_int, cell, face, kind, tinv, easi_map, easi_const, displacement, rho_g, m,
                       stream=None):
        """Nodal boundary conditions (Dirichlet / FreeSurfaceGravity faces) after `local`."""
        _check(self.lib.ssb_local_nodal_bc(self.handle, dofs, idofs, neigh_int, _ptr(cell), _ptr(face), _ptr(kind), _ptr(tinv),
                                           _ptr(easi_map), _ptr(easi_const), _ptr(displacement), _ptr(rho_g), m,
                                           _stream(stream)))

    def local_nodal_flux(self, dofs, neigh_int, cell, face, inodal, m, stream=None):
        """FaceType::Analytical faces: nodal flux of GIVEN boundary values inodal [m, 9, B2] (see boundary.py)."""
        _check(self.lib.ssb_local_nodal_flux(self.handle, dofs, neigh_int, _ptr(cell), _ptr(face), _ptr(inodal), m,
                                             _stream(stream)))

    def volume_energies(self, dofs, volume, material, n, out, stream=None):
        """Energy output of the volume: out[8] (f64 device tensor, accumulated) += kinetic / elastic / acoustic energies
        and momentum of n elements; volume [n], material [n, 3] = lambda, mu, rho (f64 device tensors)."""
        _check(self.lib.ssb_volume_energies(self.handle, dofs, _ptr(volume), _ptr(material), n, _ptr(out), _stream(stream)))

    def fault_energies(self, frictional_energy, accumulated_slip, doubled_surface_area, material_plus, material_minus, m,
                       out, plus_side_on_this_rank=None, stream=None):
        """Energy output of the fault: out[4] (f64 device tensor, accumulated) += frictional work, potency, seismic moment."""
        _check(self.lib.ssb_fault_energies(self.handle, _ptr(frictional_energy), _ptr(accumulated_slip),
                                           _ptr(doubled_surface_area), _ptr(material_plus), _ptr(material_minus),
                                           _ptr(plus_side_on_this_rank), m, _ptr(out), _stream(stream)))

    # -- set-up on the device --------------------------------------------------------------------
    def setup_cell_matrices(self, coords, neighbors, material, face_types, local_int, neigh_int, n, stream=None):
        """coords [n, 4, 3] f64, neighbors [n, 4] i64, material [*, 3] f64 (device tensors) -> local_int, neigh_int."""
        _check(self.lib.ssb_setup_cell_matrices(self.handle, _ptr(coords), _ptr(neighbors), _ptr(material), _ptr(face_types),
                                                local_int, neigh_int, n, _stream(stream)))

    def setup_dr_matrices(self, coords_plus, coords_minus, face_info, material_plus, material_minus, tinv_t, imp_and_eta,
                          flux_solver_plus, flux_solver_minus, doubled_surface_area, m, stream=None):
        _check(self.lib.ssb_setup_dr_matrices(self.handle, _ptr(coords_plus), _ptr(coords_minus), _ptr(face_info),
                                              _ptr(material_plus), _ptr(material_minus), _ptr(tinv_t), _ptr(imp_and_eta),
                                              _ptr(flux_solver_plus), _ptr(flux_solver_minus), _ptr(doubled_surface_area), m,
                                              _stream(stream)))

    # -- dynamic rupture ------------------------------------------------------------------------
    def dr_interp(self, der_plus, der_minus, tinv_t, face_info, q_plus, q_minus, coeffs, n, perm=None, stream=None):
        c = np.ascontiguousarray(coeffs, dtype=np.float64)
        assert c.shape == (self.order, self.order)
        _check(self.lib.ssb_dr_interp(self.handle, der_plus, der_minus, _ptr(tinv_t), _ptr(face_info), _ptr(q_plus),
                                      _ptr(q_minus), c.ctypes.data_as(ctypes.c_void_p), n, _ptr(perm), _stream(stream)))

    def friction(self, law, data, params, n, stream=None):
        """law: 'lsw' | 'rsfvw'; data: dict name -> torch tensor (FRICTION_DATA_FIELDS); params: SsbFrictionParams."""
        d = SsbFrictionData()
        for name in FRICTION_DATA_FIELDS:
            t = data.get(name)
            setattr(d, name, None if t is None else t.data_ptr())
        fn = self.lib.ssb_friction_lsw if law == "lsw" else self.lib.ssb_friction_rsfvw
        _check(fn(self.handle, ctypes.byref(d), ctypes.byref(params), n, _stream(stream)))

    def neigh_dr(self, dofs, godunov_ptrs, flux_solver_ptrs, side_rel, face_types, n, stream=None):
        _check(self.lib.ssb_neigh_dr(self.handle, dofs, _ptr(godunov_ptrs), _ptr(flux_solver_ptrs), _ptr(side_rel),
                                     _ptr(face_types), n, _stream(stream)))
