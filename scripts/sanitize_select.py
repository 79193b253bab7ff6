"""Node ids of the toy-shape GPU tests scripts/sanitize.sh runs under compute-sanitizer: every kernel family that
hand-rolls mbarrier / TMEM / cluster protocols, at shapes small enough for a 10-100x slowdown."""
import subprocess
import sys

WANT = [
    ("tests/test_gemm_gpu.py", ["test_gemm_bias[128-128-64]", "test_gemm_bias[256-512-256]", "test_gemm_bias[1000-136-72]",
                                "test_gemm_gate_resid_batched_views", "test_gemm_gelu_silu"]),
    ("tests/test_train_kernels_gpu.py", ["test_gemm_dgrad[1-128-128-64]", "test_gemm_dgrad[1-200-136-72]", "test_gemm_dgrad[2-300-256-512]",
                                         "test_gemm_dgrad[2-1024-1024-4096]", "test_gemm_dgrad_pitched_views_and_epilogues",
                                         "test_gemm_wgrad[1-64-128-128]", "test_gemm_wgrad[1-100-136-200]", "test_gemm_wgrad[3-150-256-384]",
                                         "test_gemm_wgrad[2-1000-1024-4608]", "test_gemm_wgrad_row_slices_of_joint_buffer",
                                         "test_attention_lse_and_backward[1-128-1]", "test_attention_lse_and_backward[2-200-2]",
                                         "test_attention_lse_and_backward[1-1000-3]", "test_gate_resid_and_backward",
                                         "test_ln_modulate_backward", "test_rmsnorm_rope_out_of_place_and_backward", "test_gelu_outer_mse",
                                         "test_adamw_matches_torch_and_clip"]),
    ("tests/test_attention_gpu.py", ["test_attention_matches_fp32_reference[1-1-1-128-128-False]",
                                     "test_attention_matches_fp32_reference[2-3-3-300-300-False]",
                                     "test_attention_matches_fp32_reference[2-2-2-640-640-False]",
                                     "test_attention_matches_fp32_reference[1-4-2-768-1000-False]",
                                     "test_attention_matches_fp32_reference[1-4-2-384-384-True]", "test_attention_peaked_softmax_rows"]),
    ("tests/test_elementwise_gpu.py", ["test_ln_modulate_matches_eager_chain[1-33-256]", "test_ln_modulate_matches_eager_chain[2-300-3072]",
                                       "test_rmsnorm_rope_matches_eager_chain", "test_euler_step_bit_exact"]),
    ("tests/test_vae_gpu.py", ["test_conv3x3_matches_torch[1-8-16-64-128-1]", "test_conv3x3_matches_torch[1-33-50-64-256-1]",
                               "test_conv3x3_matches_torch[1-32-48-128-128-2]", "test_groupnorm_silu_matches_torch_chain[1-48-32-1]"]),
]

if __name__ == "__main__":
    print(" ".join(f"{f}::{t}" for f, tests in WANT for t in tests))
