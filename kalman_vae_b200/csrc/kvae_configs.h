// kvae_configs.h — the (z_dim, a_dim, u_dim, modes) shapes the library is instantiated for.
// Each shape is built for the two dynamics variants of the reference:
//   lstm      (dyn_param.py)        : Q fixed [n,n],  C_t = sum_k alpha_k C_k
//   switching (switch_dyn_param.py) : Q_t = sum_k alpha_k Q_k, C_t = C_0
#pragma once
// X(N, P, M, K)   (a development build may override the list: -D'KVAE_FOR_EACH_SHAPE(X)=X(4,2,4,3)')
#ifndef KVAE_FOR_EACH_SHAPE
#define KVAE_FOR_EACH_SHAPE(X) \
  X(4, 2, 4, 3)                \
  X(4, 2, 4, 1)                \
  X(2, 1, 1, 1)                \
  X(8, 4, 8, 4)                \
  X(16, 8, 16, 8)
#endif
