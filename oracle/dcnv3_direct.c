/*
 * oracle/dcnv3_direct.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C + OpenMP) of the DCNv3 core in its direct,
 * pixel-space form: forward and the three analytic gradients.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product path (yolo_somi_b200/) never does.
 *
 * Semantics follow the reference (paths relative to /root/reference):
 *   output size ............ models/ops_dcnv3/src/cuda/dcnv3_cuda.cu:40-45
 *   sampling location ...... models/ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh:232-260
 *                            (point index p = i*K_h + j, i over kernel width
 *                            (outer), j over kernel height (inner); offsets are
 *                            stored (dx, dy) interleaved)
 *   range test ............. dcnv3_im2col_cuda.cuh:262-263
 *   bilinear, zero border .. dcnv3_im2col_cuda.cuh:32-80
 *   gradients .............. dcnv3_im2col_cuda.cuh:82-147
 * and were checked against the reference's own Python implementation
 * dcnv3_core_pytorch (models/ops_dcnv3/functions/dcnv3_func.py:147-188) through
 * the golden vectors in tests/golden/ (see tests/golden/make_golden.py).
 *
 * Layout: value [N,H,W,G*gc], offset [N,Ho,Wo,G*P*2], mask [N,Ho,Wo,G*P],
 * out / grad_out [N,Ho,Wo,G*gc]; all dense, channel fastest.
 *
 * Two instantiations: real = double ("truth") and real = float (the same
 * expression order a fp32 device kernel uses, so fp32 rounding is comparable).
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

typedef struct {
    int N, H, W, Ho, Wo, G, gc;
    int kh, kw, sh, sw, ph, pw, dh, dw;
    double offset_scale;
} dcnv3_geom;

int dcnv3_oracle_out_size(int in, int pad, int dil, int k, int stride) {
    /* dcnv3_cuda.cu:40-45 */
    return (in + 2 * pad - (dil * (k - 1) + 1)) / stride + 1;
}

#define REAL double
#define SUFFIX(name) name##_f64
#define FLOOR floor
#include "dcnv3_direct_impl.inc"
#undef REAL
#undef SUFFIX
#undef FLOOR

#define REAL float
#define SUFFIX(name) name##_f32
#define FLOOR floorf
#include "dcnv3_direct_impl.inc"
#undef REAL
#undef SUFFIX
#undef FLOOR
