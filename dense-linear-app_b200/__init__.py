"""dense-linear-app_b200 — B200-native tiled FP64 Cholesky.

Drop-in for the one hot path of HugoVuach/Dense-linear-app: POTRF / TRSM / SYRK / GEMM on
b x b column-major FP64 tiles of an SPD matrix (ArmoniK tile workers ``w_c_cons_v1/v2``,
Chameleon-VM ``v6_test`` / ``bench`` drivers).  Python builds the right-looking tile DAG and
calls hand-written sm_100a kernels through the C ABI of ``libchol_b200.so``
(``include/chol_b200.h``); PyTorch is used for device buffers, streams and NCCL only.

Import as ``dense_linear_app_b200`` (alias package at the repo root).
"""
__version__ = "0.1.0"
