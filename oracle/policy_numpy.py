"""Second, independent restatement of the bi-head pointer model in numpy float64 -- TEST INFRASTRUCTURE, not product.

Written from the layer list of /root/reference/ofighters/agents/qlearnIA_V2.py:123-190 and the published semantics of
the Keras layers it names -- NOT from oracle/policy_torch.py -- so that the two oracles do not share an author's slip:
everything stays NHWC like Keras, each layer is spelled out with explicit index arithmetic (shifted-slice sums for
Conv2D, a reshape-max for MaxPooling2D, an explicit source-coordinate formula for UpSampling2D) and no deep-learning
library is involved.  tests/test_policy_oracle.py checks that both restatements agree to 1e-6.

PARITY UNPINNED (see policy_torch.py): without Keras / TensorFlow and without reference weights or outputs two agreeing
restatements remove single-author risk, they do not pin the model to Keras.

Layer semantics used (Keras 2.x documentation):
  Conv2D(filters, (3,3), padding='same')   out[y,x,o] = b[o] + sum_{i,j,c} in[y+i-1, x+j-1, c] * K[i,j,c,o], zeros outside
  BatchNormalization() at inference          gamma * (x - moving_mean) / sqrt(moving_var + 0.001) + beta   (axis = -1)
  Activation('relu'), MaxPooling2D((2,2))    max over non-overlapping 2 x 2 windows (sizes are even at every level)
  Flatten() of (25,25,8) channels_last       index (h * 25 + w) * 8 + c
  Concatenate()([vector_input, flat])        the 8 vector entries FIRST (:154)
  Dense(n, activation)                       act(x @ kernel + bias), kernel (in, out)
  Reshape((25,25,1))                         index h * 25 + w
  UpSampling2D((2,2), interpolation='bilinear')
      TF2 (tf.image.resize, half_pixel_centers): src = (dst + 0.5) / 2 - 0.5, clamped to [0, n-1]; linear in each axis
      TF1 legacy (resize_bilinear, align_corners=False): src = dst / 2, upper neighbour clamped to n-1
"""
import numpy as np

BN_EPSILON = 1e-3


def conv3x3_same(x, kernel, bias):
    """x [B,H,W,Cin] float64, kernel [3,3,Cin,Cout] (HWIO), bias [Cout]."""
    B, H, W, _ = x.shape
    padded = np.zeros((B, H + 2, W + 2, x.shape[3]), dtype=np.float64)
    padded[:, 1:H + 1, 1:W + 1, :] = x
    out = np.zeros((B, H, W, kernel.shape[3]), dtype=np.float64)
    for i in range(3):
        for j in range(3):
            out += np.einsum("bhwc,co->bhwo", padded[:, i:i + H, j:j + W, :], kernel[i, j])
    return out + bias


def batchnorm_inference(x, gamma, beta, mean, var):
    return gamma * (x - mean) / np.sqrt(var + BN_EPSILON) + beta


def maxpool2(x):
    B, H, W, C = x.shape
    return x.reshape(B, H // 2, 2, W // 2, 2, C).max(axis=(2, 4))


def _resize_axis(x, axis, mode):
    n = x.shape[axis]
    dst = np.arange(2 * n, dtype=np.float64)
    if mode == "tf2":
        src = np.clip((dst + 0.5) / 2.0 - 0.5, 0.0, n - 1.0)
    elif mode == "tf1":
        src = dst / 2.0
    else:
        raise ValueError("bilinear must be 'tf2' or 'tf1'")
    lo = np.floor(src).astype(np.int64)
    hi = np.minimum(lo + 1, n - 1)
    frac = src - lo
    shape = [1] * x.ndim
    shape[axis] = 2 * n
    frac = frac.reshape(shape)
    return np.take(x, lo, axis=axis) * (1.0 - frac) + np.take(x, hi, axis=axis) * frac


def upsample2x_bilinear(x, mode="tf2"):
    return _resize_axis(_resize_axis(x, 1, mode), 2, mode)


def forward(w, image, vector, bilinear="tf2", return_intermediates=False):
    """w: dict name -> array-like in Keras layouts; image [B,400,400,2]; vector [B,8] -> act [B,2], ptr [B,400,400]."""
    W = {k: np.asarray(v, dtype=np.float64) for k, v in w.items()}
    x = np.asarray(image, dtype=np.float64)
    vec = np.asarray(vector, dtype=np.float64)
    inter = {}
    for n in range(1, 5):                                  # conv_layer{n} -> norm{n} -> relu{n} -> pool_layer{n}   (:129-147)
        x = conv3x3_same(x, W["conv%d/kernel" % n], W["conv%d/bias" % n])
        x = batchnorm_inference(x, W["norm%d/gamma" % n], W["norm%d/beta" % n], W["norm%d/mean" % n], W["norm%d/var" % n])
        x = maxpool2(np.maximum(x, 0.0))
        inter["pool%d" % n] = x
    flat = x.reshape(x.shape[0], -1)                       # Flatten(), channels_last                              (:150)
    concat = np.concatenate([vec, flat], axis=1)           # Concatenate()([vector_input, flat_layer])             (:154)
    dense1 = np.maximum(concat @ W["dense1/kernel"] + W["dense1/bias"], 0.0)
    inter["dense1"] = dense1
    dense2 = np.maximum(dense1 @ W["dense2/kernel"] + W["dense2/bias"], 0.0)
    act = dense2 @ W["output1/kernel"] + W["output1/bias"]                                                     # (:160)
    up = np.maximum(dense1 @ W["updense1/kernel"] + W["updense1/bias"], 0.0).reshape(-1, 25, 25, 1)             # (:163-164)
    for n in range(1, 5):                                  # upsampling{n} -> upconv{n} [-> upnorm{n} -> relu]     (:166-186)
        up = upsample2x_bilinear(up, bilinear)
        up = conv3x3_same(up, W["upconv%d/kernel" % n], W["upconv%d/bias" % n])
        if n < 4:
            up = batchnorm_inference(up, W["upnorm%d/gamma" % n], W["upnorm%d/beta" % n], W["upnorm%d/mean" % n], W["upnorm%d/var" % n])
            up = np.maximum(up, 0.0)
        inter["up%d" % n] = up
    ptr = up[..., 0]
    if return_intermediates:
        return act, ptr, inter
    return act, ptr


def decode(act, ptr):
    """Trainer.get_best_action (:218-220): np.argmax(act); np.unravel_index(np.argmax(ptr), ptr.shape, order='F') --
    stated with numpy itself, as the reference does."""
    out = []
    for b in range(ptr.shape[0]):
        ipointer = np.unravel_index(np.argmax(ptr[b], axis=None), ptr[b].shape[0:2], order="F")
        out.append((int(np.argmax(act[b])), (int(ipointer[0]), int(ipointer[1]))))
    return out
