"""TEST INFRASTRUCTURE: fp32 CPU restatement of the reference's ResNet policy/value forward (the checker for K4).

Follows src/networks/blocks.cpp:32-127 (layer order, kernel sizes, activations incl. tanh in the Q head, softmax axes),
src/networks/networks.cpp:71-93 / 143-168 (assembly, 32 input channels), src/networks/NNInputFeatures.cpp:65-90
(channel c of the input = bit c of the feature word; ml::unpackInput, AGNetwork.cpp:249-257) and
src/networks/NetworkDataPack.cpp:112-129 (output shapes: policy [B, H*W], value [B, 3] = win, draw, loss, q [B, H, W, 3]).

PARITY UNPINNED: the arithmetic of the reference lives in MinML, which is neither vendored nor pinned (CMakeLists.txt:8,
56-79), and no reference test evaluates a network (selfcheck.cpp:218-220 is empty). This restatement pins the CUDA kernel
to the published layer graph, not to MinML's bits. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it."""
import numpy as np
import torch
import torch.nn.functional as F


def unpack_features(features, rows, cols):
    """uint32 [B, rows*cols] -> float32 [B, 32, rows, cols] (channel c = bit c)."""
    f = torch.as_tensor(np.ascontiguousarray(features).astype(np.int64)).reshape(-1, rows, cols)
    bits = torch.arange(32, dtype=torch.int64).reshape(1, 32, 1, 1)
    return ((f.unsqueeze(1) >> bits) & 1).to(torch.float32)


def _conv(x, w, b, relu=True):
    # blob layout W[o][ky][kx][i] -> torch [o][i][ky][kx]
    wt = torch.as_tensor(w).permute(0, 3, 1, 2).contiguous()
    y = F.conv2d(x, wt, torch.as_tensor(b), padding=w.shape[1] // 2)
    return F.relu(y) if relu else y


def forward(tensors, features, rows, cols, blocks, q_head=False, activation_dtype=None):
    """Returns policy [B, rows*cols], value [B, 3], q [B, rows*cols, 3] or None, all float32.
    activation_dtype=torch.bfloat16 rounds every stored activation like the device kernel does (for error attribution)."""
    def store(t):
        return t.to(activation_dtype).to(torch.float32) if activation_dtype is not None else t

    def weights(name):
        w = torch.as_tensor(tensors[name])
        return store(w).numpy() if activation_dtype is not None and w.ndim == 4 else tensors[name]

    with torch.no_grad():
        x = unpack_features(features, rows, cols)
        x = store(_conv(x, weights("stem.w"), tensors["stem.b"]))
        for i in range(blocks):
            h = store(_conv(x, weights(f"block{i}.w1"), tensors[f"block{i}.b1"]))
            y = _conv(h, weights(f"block{i}.w2"), tensors[f"block{i}.b2"], relu=False)
            x = store(F.relu(x + y))
        n = x.shape[0]
        # policy head
        p = _conv(x, weights("policy.w"), tensors["policy.b"])
        logits = torch.einsum("bchw,oc->bohw", p, torch.as_tensor(tensors["policy.w1"])) + torch.as_tensor(tensors["policy.b1"]).reshape(1, 1, 1, 1)
        policy = torch.softmax(logits.reshape(n, -1), dim=1)
        # value head
        v = F.relu(torch.einsum("bchw,oc->bohw", x, torch.as_tensor(tensors["value.w"])) + torch.as_tensor(tensors["value.b"]).reshape(1, 4, 1, 1))
        flat = v.permute(0, 2, 3, 1).reshape(n, -1)  # NHWC flatten
        d1 = F.relu(flat @ torch.as_tensor(tensors["value.wd1"]).T + torch.as_tensor(tensors["value.bd1"]))
        value = torch.softmax(d1 @ torch.as_tensor(tensors["value.wd2"]).T + torch.as_tensor(tensors["value.bd2"]), dim=1)
        q = None
        if q_head:
            t = torch.tanh(_conv(x, weights("q.w"), tensors["q.b"], relu=False))
            ql = torch.einsum("bchw,oc->bhwo", t, torch.as_tensor(tensors["q.w1"])) + torch.as_tensor(tensors["q.b1"])
            q = torch.softmax(ql, dim=3).reshape(n, rows * cols, 3).numpy()
        return policy.numpy(), value.numpy(), q
