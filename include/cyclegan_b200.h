/*
 * cyclegan_b200.h -- C ABI of libcyclegan_b200.so (hand-written sm_100a CUDA).
 *
 * Drop-in boundary for ONE hot path of dogeplusplus/cyclegan-cat: the CycleGAN
 * training step and the generator forward.  The reference has no FFI layer of
 * its own (it is pure Python on TensorFlow/Keras); each entry point below names
 * the reference interface it replaces (paths relative to /root/reference).
 * The Python host package (cyclegan_cat_b200/) binds these with ctypes; the
 * reference-side stub a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - every function returns CG_OK (0) or a negative cg_status; it never throws
 *    or aborts across the ABI; cg_last_error() returns a thread-local message.
 *  - all pointers marked "dev" are CUDA device pointers owned by the CALLER
 *    (torch allocates them); the library owns only plan objects.
 *  - image tensors cross the boundary as float32 NHWC, exactly what the
 *    reference feeds Keras (model.py:93-106, predict.py:32).
 *  - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it.
 *  - there is no CPU fallback and no cuDNN/cuBLAS/Triton inside.
 */
#ifndef CYCLEGAN_B200_H
#define CYCLEGAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum cg_status {
    CG_OK = 0,
    CG_ERR_INVALID = -1,    /* bad argument / unsupported layer graph          */
    CG_ERR_CUDA = -2,       /* a CUDA runtime / driver call failed             */
    CG_ERR_WORKSPACE = -3,  /* caller workspace too small                      */
    CG_ERR_STATE = -4,      /* call order (e.g. backward before forward)       */
    CG_ERR_COMM = -5        /* NCCL failure                                    */
} cg_status;

/* arithmetic mode of a net (BASELINE.json north_star: bf16 mode and fp32 check mode) */
typedef enum cg_mode {
    CG_MODE_BF16 = 0,        /* NHWC bf16 activations, fp32 accumulate, tcgen05 convs  */
    CG_MODE_FP32_CHECK = 1   /* everything fp32 on CUDA cores (slow, exact twin)       */
} cg_mode;

/* Layer kinds = the Keras layers the four reference builders instantiate. */
typedef enum cg_op {
    CG_OP_CONV = 1,      /* Conv2D            unet.py:25,54,63,121  resnet.py:28,33,40,50,96,103 */
    CG_OP_CONVT = 2,     /* Conv2DTranspose   unet.py:66,76         resnet.py:57                 */
    CG_OP_INORM = 3,     /* tfa InstanceNormalization unet.py:30,56,70  resnet.py:29,34,44,51,58,98 */
    CG_OP_ACT = 4,       /* ReLU / LeakyReLU / Activation  unet.py:32,60,74,122  resnet.py:30,42,45,101 */
    CG_OP_RPAD = 5,      /* ReflectionPadding2D resnet.py:11-23                                  */
    CG_OP_ADD = 6,       /* Add               resnet.py:35                                       */
    CG_OP_CONCAT = 7,    /* Concatenate [in0, in1] on channels  unet.py:68,118                   */
    CG_OP_AVGPOOL = 8,   /* AveragePooling2D() 2x2/2   unet.py:101                               */
    CG_OP_UPSAMPLE = 9,  /* UpSampling2D() nearest x2  unet.py:109                               */
    CG_OP_BNORM = 10,    /* BatchNormalization  unet.py:27-28,57-58,71-72  resnet.py:99-100      */
    CG_OP_DROPOUT = 11   /* Dropout(0.5)      unet.py:33-34                                      */
} cg_op;

typedef enum cg_act {
    CG_ACT_NONE = 0, CG_ACT_RELU = 1, CG_ACT_LEAKY = 2, CG_ACT_TANH = 3, CG_ACT_SIGMOID = 4
} cg_act;

typedef enum cg_loss {          /* get_loss_obj, losses.py:67-81 */
    CG_LOSS_MSE = 0, CG_LOSS_MAE = 1, CG_LOSS_BCE = 2
} cg_loss;

/*
 * One node of the forward graph.  Tensor 0 is the network input (C = 3,
 * unet.py:48,92 / resnet.py:65,91); layer i produces tensor i+1; in0/in1 are
 * tensor ids.  Trainable variables are laid out in ONE flat float32 buffer in
 * Keras `trainable_variables` order: per layer [kernel, bias] or [gamma, beta].
 * Kernel layouts are TensorFlow's: Conv2D (kh,kw,Cin,Cout), Conv2DTranspose
 * (kh,kw,Cout,Cin).
 */
typedef struct cg_layer_desc {
    int32_t op;        /* cg_op                                                      */
    int32_t in0;       /* input tensor id                                            */
    int32_t in1;       /* second input (ADD, CONCAT) or -1                           */
    int32_t cin;       /* input channels (CONV/CONVT); channels otherwise            */
    int32_t cout;      /* output channels                                            */
    int32_t k;         /* square kernel size                                         */
    int32_t stride;    /* 1 or 2                                                     */
    int32_t same;      /* 1 = padding='same' (TF asymmetric rule), 0 = 'valid'       */
    int32_t has_bias;  /* use_bias                                                   */
    int32_t act;       /* cg_act (ACT layers)                                        */
    int32_t affine;    /* INORM: center=scale=True -> [gamma, beta] variables        */
    int32_t pad;       /* RPAD amount (same on H and W)                              */
    float eps;         /* INORM / BNORM epsilon (TFA and Keras default 1e-3)         */
    float slope;       /* LeakyReLU alpha                                            */
    float momentum;    /* BNORM moving-average momentum (Keras default 0.99)         */
    float rate;        /* DROPOUT rate (unet.py:34: 0.5)                             */
} cg_layer_desc;

typedef struct cg_var_info {
    int32_t layer;     /* index into the layer list                */
    int32_t role;      /* 0 kernel, 1 bias, 2 gamma, 3 beta        */
    int32_t ndim;
    int32_t shape[4];
    int64_t offset;    /* in floats, into the flat parameter buffer */
} cg_var_info;

typedef enum cg_opt {           /* get_optimizer, optimizers.py:14-21 */
    CG_OPT_ADAM = 0,            /* Adam(learning_rate, beta_1); beta_2 0.999, epsilon 1e-7 (Keras defaults)      */
    CG_OPT_SGD = 1,             /* SGD(learning_rate): p -= lr*g (momentum 0)                                     */
    CG_OPT_RMSPROP = 2,         /* RMSprop(learning_rate): rho = beta_2 (0.9), epsilon 1e-7, no momentum/centering */
    CG_OPT_ADABELIEF = 3        /* adabelief_tf AdaBeliefOptimizer(learning_rate): betas .9/.999, eps 1e-14, rectified */
} cg_opt;

typedef struct cg_adam_cfg {    /* one optimizer of optimizers.py:5-24; the slots are the trainer's m / v buffers:
                                   Adam, AdaBelief use both, RMSprop keeps its `rms` slot in v, SGD has none */
    float learning_rate, beta_1, beta_2, epsilon;
    int32_t kind;               /* cg_opt */
} cg_adam_cfg;

typedef struct cg_train_cfg {   /* configs/cycle.yaml:36-41 + training_config.yaml:4-11 */
    int32_t loss;               /* cg_loss                                      */
    float w_cycle, w_identity, w_generator, w_discriminator;
    cg_adam_cfg adam[4];        /* order: g_AB, g_BA, d_A, d_B (model.py:68-71) */
} cg_train_cfg;

typedef struct cg_net_s* cg_net_t;
typedef struct cg_trainer_s* cg_trainer_t;

/* ---- library ------------------------------------------------------------ */
int cg_init(int device);                 /* replaces train.py:36-43 (device selection); checks sm_100 */
const char* cg_last_error(void);
int cg_version(void);
int cg_abi_sizeof(int which);            /* sizeof of 0 cg_layer_desc, 1 cg_var_info, 2 cg_train_cfg, 3 cg_adam_cfg: lets a
                                            binding check its struct mirrors before the first real call */

/* ---- model builder: replaces keras.Model(inputs, outputs) at unet.py:78,123 / resnet.py:85,105 */
int cg_net_create(const cg_layer_desc* layers, int n_layers, int mode, cg_net_t* out);
void cg_net_destroy(cg_net_t net);
int cg_net_param_floats(cg_net_t net, int64_t* n_floats);
int cg_net_var_count(cg_net_t net, int* n_vars);                   /* len(model.trainable_variables) */
int cg_net_var_info(cg_net_t net, int i, cg_var_info* out);
int cg_net_out_shape(cg_net_t net, int N, int H, int W, int out_nhwc[4]);
int cg_net_workspace_bytes(cg_net_t net, int N, int H, int W, int need_backward, size_t* bytes);
/* Non-trainable state = the BatchNormalization moving statistics (Keras `non_trainable_variables` order: per BNORM
 * layer [moving_mean[C], moving_variance[C]]); 0 floats for instance-norm nets.  The buffer is caller-owned. */
int cg_net_state_floats(cg_net_t net, int64_t* n_floats);
int cg_net_bind_state(cg_net_t net, float* state_dev);
/* Keras `model(x, training=...)`: selects batch statistics + moving-average update (BNORM) and active DROPOUT for the
 * following cg_net_forward calls on this handle; default 0 (inference), like Keras.  Ignored by instance norm. */
int cg_net_set_training(cg_net_t net, int training);
/* DROPOUT masks are a counter-based hash of (seed, call counter, layer, element): reproducible, and restated bit for
 * bit by the oracle (TensorFlow's own random stream cannot be reproduced). */
int cg_net_set_seed(cg_net_t net, uint64_t seed);

/* model(x) -- Keras Model.__call__ (predict.py:32,35; model.py:93-106; unittests/test_*.py) */
int cg_net_forward(cg_net_t net, const float* params_dev, const float* x_dev, float* y_dev,
                   void* workspace_dev, size_t workspace_bytes, int N, int H, int W,
                   int need_backward, void* stream);
/* tape.gradient(loss, model.trainable_variables) for one model call (model.py:143-147):
 * dy is dLoss/dy of the preceding cg_net_forward on the same workspace. */
int cg_net_backward(cg_net_t net, const float* params_dev, const float* dy_dev, float* dx_dev_or_null,
                    float* grads_dev, int accumulate, void* workspace_dev, size_t workspace_bytes,
                    void* stream);

/* Intermediate tensor `tensor` (0 = input, i+1 = output of layer i) of the last cg_net_forward on this handle, as float32
 * NHWC; shape4 (nullable) receives its shape.  With out_dev == NULL it only queries: returns 0 when the last forward
 * materialised the tensor, 1 when a fusion skipped it (e.g. an InstanceNormalization folded into its ReLU / reflection
 * pad).  Replaces `keras.Model(inputs, layer.output)` probes; the parity tests use it to compare every layer with the
 * oracle on identical inputs (tests/test_gpu_layerwise.py). */
int cg_net_fetch_tensor(cg_net_t net, int tensor, float* out_dev, int shape4[4], void* stream);

/* ---- trainer: replaces CycleGan.train_step / validate_step (model.py:91-154) */
int cg_trainer_create(cg_net_t g_AB, cg_net_t g_BA, cg_net_t d_A, cg_net_t d_B,
                      const cg_train_cfg* cfg, cg_trainer_t* out);
void cg_trainer_destroy(cg_trainer_t tr);
int cg_trainer_workspace_bytes(cg_trainer_t tr, int B, int H, int W, size_t* bytes);
/* params/grads/adam_m/adam_v: four flat float32 device buffers each (g_AB, g_BA, d_A, d_B). */
int cg_trainer_bind(cg_trainer_t tr, float* const params_dev[4], float* const grads_dev[4],
                    float* const adam_m_dev[4], float* const adam_v_dev[4],
                    void* workspace_dev, size_t workspace_bytes);
/* metrics6_dev: gAB_loss, gBA_loss, dA_loss, dB_loss, dA_acc, dB_acc (model.py:126-133), device floats */
int cg_validate_step(cg_trainer_t tr, const float* real_a_dev, const float* real_b_dev,
                     int B, int H, int W, float* metrics6_dev, void* stream);
int cg_train_step(cg_trainer_t tr, const float* real_a_dev, const float* real_b_dev,
                  int B, int H, int W, float* metrics6_dev, void* stream);
/* the two halves of train_step, exposed for tests (per-layer gradient parity) */
int cg_trainer_compute_gradients(cg_trainer_t tr, const float* real_a_dev, const float* real_b_dev,
                                 int B, int H, int W, float* metrics6_dev, void* stream);
int cg_trainer_apply_gradients(cg_trainer_t tr, void* stream);      /* 4x optimizer.apply_gradients, model.py:149-153 */
/* optimizer.apply_gradients(zip(grads, variables)) on its own (model.py:149-153, and the zero-gradient step of
 * load_optimizer, model.py:359-362): ONE fused launch over a flat float32 range.  `iterations` is optimizer.iterations
 * before the step (the caller increments it); slot_m / slot_v are the Keras slots (Adam, AdaBelief: m and v; RMSprop:
 * rms in slot_v; SGD: none, both may be NULL). */
int cg_optimizer_apply(const cg_adam_cfg* cfg, float* params_dev, const float* grads_dev, float* slot_m_dev,
                       float* slot_v_dev, size_t n, int64_t iterations, void* stream);
int cg_trainer_get_iterations(cg_trainer_t tr, int64_t iters[4]);   /* optimizer.iterations (model.py:314-315)  */
int cg_trainer_set_iterations(cg_trainer_t tr, const int64_t iters[4]);
/* The trainer keeps the plans (buffer layout, TMA descriptors, captured CUDA graphs) of the last few batch shapes it has
 * seen, so the ragged last batch of an epoch (model.py:197, `dataset.batch(batch_size)` without drop_remainder) switches
 * plans instead of re-planning.  plans_built = layouts computed so far, parked = plans kept besides the current one. */
int cg_trainer_plan_count(cg_trainer_t tr, int64_t* plans_built, int* parked);
/* pointer to an image the last step produced, for tests: 0 fake_b 1 same_b 2 fake_a 3 same_a 4 cycled_a 5 cycled_b */
int cg_trainer_fetch_image(cg_trainer_t tr, int which, float* out_dev, void* stream);

/* the same probe as cg_net_fetch_tensor for one of the six model calls of the last step:
 * call 0 g_AB([a;b]) 1 g_BA([b;a]) 2 g_BA(fake_b) 3 g_AB(fake_a) 4 d_A([a;fake_a]) 5 d_B([b;fake_b])  (model.py:93-106).
 * Calls 1 and 2 are executed as ONE g_BA launch sequence over [fake_b; b; a]; the probe returns each call's sample range. */
int cg_trainer_fetch_tensor(cg_trainer_t tr, int call, int tensor, float* out_dev, int shape4[4], void* stream);

/* ---- input pipeline (transform/data_load.py:20-34, predict.py:20-27), all HBM-bound streaming kernels ---------- */
/* normalize (data_load.py:31-34): dst = float32(src) / 127.5 - 1 */
int cg_normalize_u8(const uint8_t* src_dev, float* dst_dev, size_t n, void* stream);
/* postprocess_prediction (predict.py:26-27): dst = uint8((src + 1) * 127.5), truncating, clamped to [0, 255] */
int cg_postprocess_u8(const float* src_dev, uint8_t* dst_dev, size_t n, void* stream);
/* tf.image.resize(x, [Ho, Wo]) (data_load.py:23,41): bilinear, half-pixel centres, no antialiasing; float32 NHWC */
int cg_resize_bilinear(const float* src_dev, int N, int H, int W, int C, float* dst_dev, int Ho, int Wo, void* stream);
/* random_jitter (data_load.py:21-27) as ONE kernel: resize to [Hr, Wr] (bilinear, as above), crop [Ho, Wo] at
 * (oy[n], ox[n]) and mirror horizontally where flip[n] != 0, without materialising the resized image.  The caller
 * draws the random offsets / flips (tf.image.random_crop / random_flip_left_right cannot be reproduced bit for bit). */
int cg_resize_crop_flip(const float* src_dev, int N, int H, int W, int C, int Hr, int Wr, float* dst_dev, int Ho, int Wo,
                        const int32_t* oy_dev, const int32_t* ox_dev, const int32_t* flip_dev, void* stream);

/* ---- data parallel (new functionality; the reference is single-device, train.py:36-43) */
int cg_comm_unique_id(char id_out[128]);
int cg_trainer_comm_init(cg_trainer_t tr, const char id[128], int rank, int world);

/* ---- instrumentation for bench.py: CUDA-event timing of one kernel family */
int cg_prof_enable(int enable);                 /* brackets every tensor-core conv launch with events; 0 = off, 1 = on,
                                                   n > 1 = on with n event pairs created up front,
                                                   -1 = pause (keep the records for cg_prof_read) */
int cg_prof_read(double* total_ms, int64_t* launches, double* total_flops);  /* syncs, resets */
int cg_launch_count(int64_t* launches, int reset);  /* kernels launched by this library */

#ifdef __cplusplus
}
#endif
#endif /* CYCLEGAN_B200_H */
