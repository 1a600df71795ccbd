"""ctypes binding of libapex_b200.so (include/apex_cassie.h).  No fallback: a missing library is an error."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# APEX_B200_LIB: developer override used by tools/build_variant.sh to A/B kernel builds on the GPU box
_LIB_PATH = os.environ.get("APEX_B200_LIB") or os.path.join(_HERE, "libapex_b200.so")
_lib = None


class ApexLibraryError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ApexLibraryError(
                f"{_LIB_PATH} is missing: build it with `python -m apex_b200.build` (nvcc, sm_100a). "
                "apex_b200 has no CPU or PyTorch fallback for the environment step.")
        L = C.CDLL(_LIB_PATH)
        vp, ip, i, u = C.c_void_p, C.c_void_p, C.c_int, C.c_uint
        L.apex_cassie_state_words.restype = i
        L.apex_cassie_istate_words.restype = i
        L.apex_cassie_layout.argtypes = [C.c_char_p]
        L.apex_cassie_layout.restype = i
        L.apex_cassie_env_init.argtypes = [i, vp, ip, i, u, i, i, vp]
        L.apex_cassie_env_reset.argtypes = [i, vp, ip, i, vp, vp]
        L.apex_cassie_env_step.argtypes = [i, vp, ip, i, vp, vp, vp, ip, vp, i, vp]
        L.apex_cassie_env_reset_for_test.argtypes = [i, vp, ip, i, vp, vp, i, vp]
        L.apex_cassie_env_reset_for_test.restype = i
        L.apex_cassie_mj_step.argtypes = [i, vp, ip, i, i, vp]
        for f in (L.apex_cassie_env_init, L.apex_cassie_env_reset, L.apex_cassie_env_step, L.apex_cassie_mj_step):
            f.restype = i
        L.apex_cassietraj_env_init.argtypes = [i, vp, ip, i, u, i, i, vp]
        L.apex_cassietraj_env_reset.argtypes = [i, vp, ip, i, vp, vp, i, i, vp]
        L.apex_cassietraj_env_step.argtypes = [i, vp, ip, i, vp, vp, vp, ip, vp, i, vp, vp, i, i, vp]
        for f in (L.apex_cassietraj_env_init, L.apex_cassietraj_env_reset, L.apex_cassietraj_env_step):
            f.restype = i
        L.apex_cassie_env_order.argtypes = [ip, i, ip, vp]
        L.apex_cassie_env_step_ordered.argtypes = [i, vp, ip, i, vp, vp, vp, ip, vp, i, vp, vp, i, i, vp, vp]
        L.apex_cassie_env_order.restype = L.apex_cassie_env_step_ordered.restype = i
        L.apex_cassie_set_warps_per_cta.argtypes = [i]
        L.apex_cassie_set_warps_per_cta.restype = None
        fl, lng = C.c_float, C.c_long
        L.apex_mlp_forward.argtypes = [vp, i, i, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.apex_mlp_backward.argtypes = [vp, i, i, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.apex_prepare_obs.argtypes = [vp, vp, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.apex_gaussian_sample.argtypes = [vp, vp, fl, i, i, u, u, u, vp, vp, vp]
        L.apex_ppo_loss.argtypes = [i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, fl, fl, vp, vp, vp, vp, vp, vp, vp]
        L.apex_grad_sumsq.argtypes = [vp, i, vp, vp]
        L.apex_adam_step.argtypes = [vp, vp, vp, vp, i, vp, fl, fl, fl, fl, fl, fl, i, vp]
        L.apex_gae_scan.argtypes = [i, i, vp, vp, vp, vp, vp, fl, fl, vp, vp, vp]
        L.apex_moments.argtypes = [vp, lng, vp, vp]
        L.apex_normalize.argtypes = [vp, lng, vp, fl, vp]
        L.apex_cassie_env_step_masked.argtypes = [i, vp, ip, i, vp, vp, vp, ip, vp, i, vp, vp]
        L.apex_cassie_env_step_masked.restype = i
        L.apex_ars_policy.argtypes = [vp, i, i, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.apex_ars_policy.restype = i
        L.apex_ars_update.argtypes = [vp, i, vp, vp, vp, i, fl, vp]
        L.apex_ars_update.restype = i
        L.apex_mlp_backward_dx.argtypes = [vp, i, i, i, i] + [vp] * 9 + [i] + [vp] * 7
        L.apex_replay_gather.argtypes = [vp, vp, i, i, i, vp, vp, vp, vp, vp, vp]
        L.apex_td3_action.argtypes = [vp, vp, vp, i, i, i, fl, fl, fl, u, u, vp, vp, vp]
        L.apex_td3_critic_loss.argtypes = [i, vp, vp, vp, vp, vp, vp, fl, vp, vp, vp, vp]
        L.apex_td3_actor_grad.argtypes = [i, i, i, vp, vp, fl, vp, vp]
        L.apex_polyak.argtypes = [vp, vp, i, fl, vp]
        for f in (L.apex_mlp_backward_dx, L.apex_replay_gather, L.apex_td3_action, L.apex_td3_critic_loss, L.apex_td3_actor_grad,
                  L.apex_polyak):
            f.restype = i
        L.apex_tc_linear_forward.argtypes = [vp, i, i, vp, vp, i, i, vp, vp]
        L.apex_mlp_forward_bf16.argtypes = [vp, i, i, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_long, vp]
        L.apex_tc_linear_tiled.argtypes = [vp, i, i, vp, vp, vp, i, i, vp, vp]
        L.apex_mlp_bf16_scratch_bytes.argtypes = [i, i]
        L.apex_mlp_bf16_scratch_bytes.restype = C.c_long
        L.apex_tc_linear_forward.restype = L.apex_mlp_forward_bf16.restype = L.apex_tc_linear_tiled.restype = i
        L.apex_set_tc_persistent.argtypes = [i]
        L.apex_set_tc_persistent.restype = None
        L.apex_set_gemm_large_tiles.argtypes = [i]
        L.apex_set_gemm_large_tiles.restype = None
        L.apex_gaussian_sample_dev.argtypes = [vp, vp, vp, i, i, u, u, vp, vp, vp]
        L.apex_gaussian_sample_dev.restype = i
        L.apex_adam_step_dev.argtypes = [vp, vp, vp, vp, i, vp, fl, fl, fl, fl, fl, fl, vp, vp]
        L.apex_td3_action_dev.argtypes = [vp, vp, i, i, i, fl, fl, fl, u, vp, vp, vp, vp]
        L.apex_replay_sample.argtypes = [vp, i, vp, u, vp, vp]
        L.apex_counter_add.argtypes = [vp, i, vp]
        L.apex_adam_step_dev.restype = L.apex_td3_action_dev.restype = L.apex_replay_sample.restype = L.apex_counter_add.restype = i
        L.apex_set_head_kernels.argtypes = [i]
        L.apex_set_head_kernels.restype = None
        L.apex_set_tc_mode.argtypes = [i]
        L.apex_set_tc_mode.restype = None
        L.apex_get_tc_mode.restype = i
        L.apex_set_tc_min_rows.argtypes = [i]
        L.apex_set_tc_min_rows.restype = None
        L.apex_tc3_linear.argtypes = [vp, lng, i, i, vp, lng, lng, vp, i, vp, lng, vp, lng, i, vp]
        L.apex_tc3_outer.argtypes = [vp, lng, vp, lng, i, lng, vp, lng, i, i, vp]
        L.apex_tc3_linear.restype = L.apex_tc3_outer.restype = i
        L.apex_col_moments.argtypes = [vp, i, i, vp, vp]
        L.apex_col_moments.restype = i
        for f in (L.apex_mlp_forward, L.apex_mlp_backward, L.apex_prepare_obs, L.apex_gaussian_sample, L.apex_ppo_loss,
                  L.apex_grad_sumsq, L.apex_adam_step, L.apex_gae_scan, L.apex_moments, L.apex_normalize):
            f.restype = i
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        raise ApexLibraryError(f"{what} failed with code {rc}")


def layout(name):
    off = lib().apex_cassie_layout(name.encode())
    if off < 0:
        raise KeyError(name)
    return off
