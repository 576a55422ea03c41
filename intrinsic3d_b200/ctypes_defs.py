"""ctypes mirrors of include/i3d_types.h (I3DParams / I3DIterInfo / I3DLightingParams / I3DLightingInfo).

Field order and types must match the C header exactly; tests/test_abi.py checks
sizeof() against the compiled library.
"""
import ctypes as C

NUM_COST_TYPES = 4
MAX_OBS = 8
EG_COLS = 29
MAX_LM_STEPS = 64


class I3DParams(C.Structure):
    _fields_ = [
        ("lambda_", C.c_double * NUM_COST_TYPES),
        ("use_er", C.c_int32),
        ("use_es", C.c_int32),
        ("use_ea", C.c_int32),
        ("fix_all_albedo", C.c_int32),
        ("thres_shell", C.c_double),
        ("occlusion_distance", C.c_float),
        ("num_observations", C.c_int32),
        ("lm_steps", C.c_int32),
        ("fix_poses", C.c_int32),
        ("fix_intrinsics", C.c_int32),
        ("fix_distortion", C.c_int32),
        ("initial_trust_region_radius", C.c_double),
        ("max_trust_region_radius", C.c_double),
        ("min_trust_region_radius", C.c_double),
        ("min_relative_decrease", C.c_double),
        ("min_lm_diagonal", C.c_double),
        ("max_lm_diagonal", C.c_double),
        ("eta", C.c_double),
        ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double),
        ("parameter_tolerance", C.c_double),
        ("max_linear_solver_iterations", C.c_int32),
        ("min_linear_solver_iterations", C.c_int32),
        ("residual_reset_period", C.c_int32),
        ("max_consecutive_invalid_steps", C.c_int32),
        ("forced_cg_iterations", C.c_int32),
        ("build_only", C.c_int32),
    ]


class I3DIterInfo(C.Structure):
    _fields_ = [
        ("num_voxels", C.c_int64),
        ("num_active", C.c_int64),
        ("num_free_sdf", C.c_int64),
        ("num_free_albedo", C.c_int64),
        ("num_parameters", C.c_int64),
        ("type_residuals", C.c_int64 * NUM_COST_TYPES),
        ("type_sum_weights", C.c_double * NUM_COST_TYPES),
        ("type_weights", C.c_double * NUM_COST_TYPES),
        ("type_costs", C.c_double * NUM_COST_TYPES),
        ("cost_initial", C.c_double),
        ("cost_final", C.c_double),
        ("trust_region_radius", C.c_double),
        ("lm_iterations", C.c_int32),
        ("step_accepted", C.c_int32),
        ("termination", C.c_int32),
        ("cg_iterations_total", C.c_int32),
        ("cg_iterations", C.c_int32 * MAX_LM_STEPS),
        ("model_cost_change", C.c_double * MAX_LM_STEPS),
        ("candidate_cost", C.c_double * MAX_LM_STEPS),
        ("relative_decrease", C.c_double * MAX_LM_STEPS),
        ("step_norm", C.c_double),
        ("time_add", C.c_double),
        ("time_build", C.c_double),
        ("time_solve", C.c_double),
    ]

    def as_dict(self):
        out = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out


def default_params() -> I3DParams:
    """Defaults = data/intrinsic3d.yml (first outer iteration) + Ceres 2.1.0 solver defaults."""
    p = I3DParams()
    p.lambda_[0], p.lambda_[1], p.lambda_[2], p.lambda_[3] = 0.2, 80.0, 120.0, 0.1
    p.use_er = p.use_es = p.use_ea = 1
    p.fix_all_albedo = 0
    p.thres_shell = 0.0
    p.occlusion_distance = 0.02
    p.num_observations = 5
    p.lm_steps = 50
    p.initial_trust_region_radius = 1e4
    p.max_trust_region_radius = 1e16
    p.min_trust_region_radius = 1e-32
    p.min_relative_decrease = 1e-3
    p.min_lm_diagonal = 1e-6
    p.max_lm_diagonal = 1e32
    p.eta = 0.1
    p.function_tolerance = 1e-6
    p.gradient_tolerance = 1e-10
    p.parameter_tolerance = 1e-8
    p.max_linear_solver_iterations = 500
    p.min_linear_solver_iterations = 0
    p.residual_reset_period = 10
    p.max_consecutive_invalid_steps = 5
    return p


class _Dictable:
    def as_dict(self):
        out = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out


class I3DLightingParams(C.Structure, _Dictable):
    _fields_ = [
        ("subvolume_size", C.c_float),
        ("weighted", C.c_int32),
        ("lambda_reg", C.c_double),
        ("thres_shell", C.c_double),
        ("max_iterations", C.c_int32),
        ("max_linear_solver_iterations", C.c_int32),
        ("min_linear_solver_iterations", C.c_int32),
        ("residual_reset_period", C.c_int32),
        ("max_consecutive_invalid_steps", C.c_int32),
        ("reserved", C.c_int32),
        ("initial_trust_region_radius", C.c_double),
        ("max_trust_region_radius", C.c_double),
        ("min_trust_region_radius", C.c_double),
        ("min_relative_decrease", C.c_double),
        ("min_lm_diagonal", C.c_double),
        ("max_lm_diagonal", C.c_double),
        ("eta", C.c_double),
        ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double),
        ("parameter_tolerance", C.c_double),
    ]


class I3DLightingInfo(C.Structure, _Dictable):
    _fields_ = [
        ("num_subvolumes", C.c_int64),
        ("num_data_rows", C.c_int64),
        ("num_reg_pairs", C.c_int64),
        ("sum_data_weights", C.c_double),
        ("cost_initial", C.c_double),
        ("cost_final", C.c_double),
        ("trust_region_radius", C.c_double),
        ("lm_iterations", C.c_int32),
        ("num_successful_steps", C.c_int32),
        ("cg_iterations_total", C.c_int32),
        ("termination", C.c_int32),
        ("usable", C.c_int32),
        ("reserved", C.c_int32),
        ("time_accumulate", C.c_double),
        ("time_solve", C.c_double),
        ("time_interpolate", C.c_double),
    ]
