"""Shared configs / helpers for the parity tests."""
import numpy as np

UNET_G = dict(type="unet_generator", filters=[16, 32, 64, 128], kernels=[4, 4, 4, 4], output_channels=3,
              expansion="upsample", normalization="instancenorm", dropout=False, final_activation="tanh")  # cycle.yaml:5-20
UNET_D = dict(type="unet_generator", filters=[16, 32, 64], kernels=[7, 5, 3], output_channels=1,
              expansion="upsample", normalization="instancenorm", dropout=False, final_activation="sigmoid")  # cycle.yaml:22-35
SIMPLE_D3 = dict(type="simple_discriminator", filters=[64, 128, 256], kernels=[4, 4, 4], normalization="instancenorm")
SIMPLE_D4 = dict(type="simple_discriminator", filters=[64, 128, 256, 512], kernels=[4, 4, 4, 4],
                 normalization="instancenorm")
RESNET64 = dict(type="resnet_generator", filters=64)
# reference unit-test fixtures (unittests/test_unet.py:8-20, test_resnet.py:7-19)
FIX_UNET = dict(type="strided_unet", filters=[8, 8, 8], kernels=[4, 4, 4], output_channels=3, expansion="upsample",
                normalization="instancenorm", dropout=False, final_activation="tanh")
FIX_RESNET = dict(type="resnet_generator", filters=16)
FIX_SIMPLE = dict(type="simple_discriminator", filters=[8, 16, 32], kernels=[4, 4, 4], normalization="instancenorm")
# small but structurally complete nets for fast gradient checks
SMALL_RESNET = dict(type="resnet_generator", filters=8)
SMALL_STRIDED = dict(type="strided_unet", filters=[8, 16, 16], kernels=[4, 3, 4], output_channels=3,
                     normalization="instancenorm", final_activation="tanh")
SMALL_UNET = dict(type="unet_generator", filters=[8, 16, 16], kernels=[4, 5, 3], output_channels=3,
                  expansion="upsample", normalization="instancenorm", dropout=False, final_activation="tanh")
SMALL_UNET_D = dict(type="unet_generator", filters=[8, 8], kernels=[7, 3], output_channels=1,
                    expansion="upsample", normalization="instancenorm", dropout=False, final_activation="sigmoid")
SMALL_SIMPLE = dict(type="simple_discriminator", filters=[8, 16], kernels=[4, 4], normalization="instancenorm")

# optional config paths: BatchNormalization / Dropout variants of the small nets
BN_STRIDED = dict(SMALL_STRIDED, normalization="batchnorm")
BN_UNET = dict(SMALL_UNET, normalization="batchnorm")
BN_SIMPLE = dict(SMALL_SIMPLE, normalization="batchnorm")
DROP_UNET = dict(SMALL_UNET, dropout=True)
BN_DROP_UNET = dict(SMALL_UNET, normalization="batchnorm", dropout=True)
NONORM_UNET = dict(SMALL_UNET, normalization="none")        # unet.py:27-30: neither branch -> no norm layer

LOSS_WEIGHTS = dict(cycle=2.0, identity=0.5, generator=1.0, discriminator=0.5)   # cycle.yaml:36-41
ADAM = dict(name="adam", learning_rate=2e-4, beta_1=0.5)                         # training_config.yaml:4-11


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def model_config(gen, disc, loss="mse"):
    from cyclegan_cat_b200.model_processing.load_model import Bunch
    return Bunch(name="model", new=True, location="/tmp/cg_b200_models", generator=dict(gen),
                 discriminator=dict(disc), loss=loss, loss_weights=dict(LOSS_WEIGHTS))


def train_config(batch_size=1, g_opt=None, d_opt=None):
    from cyclegan_cat_b200.model_processing.load_model import Bunch
    return Bunch(epochs=1, batch_size=batch_size, image_size=128, g_opt=dict(g_opt or ADAM), d_opt=dict(d_opt or ADAM),
                 summary=dict(samples=1, images=5, model=20))
