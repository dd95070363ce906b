/* Diagnostics-only entry points: a naive comparator convolution and the UMMA descriptor / MMA-rate probes used while
 * bringing the kernels up (tests/gpu_first_light.py).  They are NOT part of the product library: libunet_b200.so does
 * not export them; `make libunet_b200_diag.so` (-DUB_DIAGNOSTICS) builds a library that has them in addition. */
#ifndef UNET_B200_DIAG_H_
#define UNET_B200_DIAG_H_
#include "unet_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

int ub_naive_conv3x3(const void* x_bf16, const void* w_fwd, const float* bias, float* y_f32, int n, int h, int w_,
                     int cin, int cout, int relu, void* stream);
/* KW, KH, stages, items, splits, chunks per split, BN, grid of the last plan of the persistent wgrad kernel */
void ub_last_wgrad_plan(int* out8);
/* Runs the descriptor-semantics probes; writes up to `cap` floats of results to host `out`. */
int ub_probe_run(int which, float* out_host, int cap);

#ifdef __cplusplus
}
#endif
#endif /* UNET_B200_DIAG_H_ */
