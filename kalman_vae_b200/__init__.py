"""kalman_vae_b200 — B200-native (sm_100a) Kalman filter / RTS smoother / ELBO / explicit adjoint behind
the call signatures of rodrigo-paganini/kalman-vae's `kvae.kalman` package.

    from kalman_vae_b200 import KalmanFilter, DynamicsParameter, SwitchingDynamicsParameter

The arithmetic lives in libkvae_kalman.so (hand-written CUDA, C ABI in include/kvae_kalman.h); there is
no CPU implementation and no fallback.
"""
from .dyn_param import DynamicsParameter, SwitchingDynamicsParameter  # noqa: F401
from .kalman_filter import KalmanFilter  # noqa: F401

__all__ = ["KalmanFilter", "DynamicsParameter", "SwitchingDynamicsParameter"]
