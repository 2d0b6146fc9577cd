"""Network blob: the weights of a ResnetPV / ResnetPVQ (src/networks/networks.cpp:71-93, 143-168; layers from
src/networks/blocks.cpp:32-127) after inference folding (AGNetwork::optimize, AGNetwork.cpp:136-149: batch-norm folded
into the preceding conv/dense, so every layer is weights + bias). All fp32, little endian, tensors in this order:

  stem      W[F][5][5][32]  b[F]                      conv5x5 + ReLU           (createInputBlock)
  block i   W1[F][3][3][F] b1[F]  W2[F][3][3][F] b2[F]  x = relu(x + conv2(relu(conv1(x))))   (createResidualBlock)
  policy    Wp[F][3][3][F] bp[F]  wp1[1][F] bp1[1]      conv3x3 + ReLU, conv1x1, softmax over the board  (createPolicyHead)
  value     Wv[4][F] bv[4]  Wd1[D][H*W*4] bd1[D]  Wd2[3][D] bd2[3]   conv1x1+ReLU, dense+ReLU, dense, softmax; D = min(256, 2F);
            dense input index = (row*W + col)*4 + channel (NHWC flatten)                              (createValueHead)
  q (pvq)   Wq[F][3][3][F] bq[F]  wq1[3][F] bq1[3]      conv3x3 + tanh, conv1x1, softmax over 3        (createActionValuesHead)

Convolutions are cross-correlations with "same" zero padding: out[y,x,o] = sum W[o,ky,kx,i] * in[y+ky-r, x+kx-r, i].
The file format of the reference's networks is MinML's and is not visible in the reference tree (SURVEY.md §7 hard part 6);
this blob is the documented boundary instead."""
import numpy as np

INPUT_CHANNELS = 32


def tensor_shapes(rows, cols, blocks, filters, q_head):
    f, d = filters, min(256, 2 * filters)
    shapes = [("stem.w", (f, 5, 5, INPUT_CHANNELS)), ("stem.b", (f,))]
    for i in range(blocks):
        shapes += [(f"block{i}.w1", (f, 3, 3, f)), (f"block{i}.b1", (f,)), (f"block{i}.w2", (f, 3, 3, f)), (f"block{i}.b2", (f,))]
    shapes += [("policy.w", (f, 3, 3, f)), ("policy.b", (f,)), ("policy.w1", (1, f)), ("policy.b1", (1,))]
    shapes += [("value.w", (4, f)), ("value.b", (4,)), ("value.wd1", (d, rows * cols * 4)), ("value.bd1", (d,)), ("value.wd2", (3, d)), ("value.bd2", (3,))]
    if q_head:
        shapes += [("q.w", (f, 3, 3, f)), ("q.b", (f,)), ("q.w1", (3, f)), ("q.b1", (3,))]
    return shapes


def blob_size(rows, cols, blocks, filters, q_head):
    return 4 * sum(int(np.prod(s)) for _, s in tensor_shapes(rows, cols, blocks, filters, q_head))


def random_tensors(rows, cols, blocks, filters, q_head, seed=1234):
    """He-normal conv/dense weights, small random biases (synthetic weights, same blob for CPU and GPU)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in tensor_shapes(rows, cols, blocks, filters, q_head):
        if len(shape) == 1:
            out[name] = (0.05 * rng.standard_normal(shape)).astype(np.float32)
        else:
            fan_in = int(np.prod(shape[1:]))
            scale = np.sqrt(2.0 / fan_in)
            if name.endswith(".w2"):
                scale *= 0.25  # keeps the residual stream bounded over 20 blocks
            out[name] = (scale * rng.standard_normal(shape)).astype(np.float32)
    return out


def pack(tensors, rows, cols, blocks, filters, q_head):
    parts = []
    for name, shape in tensor_shapes(rows, cols, blocks, filters, q_head):
        t = np.ascontiguousarray(tensors[name], np.float32)
        assert t.shape == tuple(shape), (name, t.shape, shape)
        parts.append(t.reshape(-1))
    return np.concatenate(parts)


def unpack(blob, rows, cols, blocks, filters, q_head):
    blob = np.asarray(blob, np.float32).reshape(-1)
    out, off = {}, 0
    for name, shape in tensor_shapes(rows, cols, blocks, filters, q_head):
        n = int(np.prod(shape))
        out[name] = blob[off:off + n].reshape(shape)
        off += n
    assert off == blob.size
    return out
