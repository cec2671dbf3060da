// wunet_train.cuh — interface of the fp32 training step (implemented in wunet_train.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace wunet {

// Parameters of the 2n+1 conv blocks and of the head, in the reference's own layouts, read straight from the torch tensors.
struct TrainParams {
    const float *const *conv_w;   // [2n+1] -> [Cout][Cin][K]
    const float *const *conv_b;   // [2n+1] -> [Cout]
    const float *const *bn_w;     // [2n+1] -> [Cout]   BatchNorm1d.weight
    const float *const *bn_b;     // [2n+1] -> [Cout]   BatchNorm1d.bias
    float *const *bn_mean;        // [2n+1] -> [Cout]   running_mean, UPDATED in place by the forward
    float *const *bn_var;         // [2n+1] -> [Cout]   running_var,  UPDATED in place by the forward
    const float *out_w;           // [1][C+1][1]
    const float *out_b;           // [1]
};
// Gradient buffers, same layouts as the parameters they belong to (fully overwritten).
struct TrainGrads {
    float *const *conv_w, *const *conv_b, *const *bn_w, *const *bn_b;
    float *out_w, *out_b;
};

const char *train_error();
size_t train_workspace_bytes(int n_layers, int ci, int B, int T);
// y = model(x) with BatchNorm in training mode; keeps every pre-BN output and activation in `workspace` for the backward
int train_forward(int n_layers, int ci, const float *x, float *y, int B, int T, const TrainParams &P, float momentum,
                  void *workspace, cudaStream_t stream);
// gradients of every parameter for the upstream gradient dy of the output; `workspace` as left by train_forward.
// part: -1 = the whole backward; 0 = head + decoder blocks (their gradients are complete afterwards: a data-parallel caller
// starts reducing them while part 1 runs); 1 = middle + encoder blocks (must follow part 0 on the same stream).
int train_backward(int n_layers, int ci, const float *x, const float *y, const float *dy, int B, int T, const TrainParams &P,
                   const TrainGrads &G, void *workspace, cudaStream_t stream, int part = -1);

}  // namespace wunet
