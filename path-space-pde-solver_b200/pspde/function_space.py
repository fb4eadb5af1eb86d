"""Function-space modules with the reference's interface (function_space.py): DenseNet (:116-140),
MySequential (:177-195), SingleParam (:6-21).

Parameter names, shapes, registration order, initialisation and RNG draw order are those of the reference, so
state_dicts are interchangeable and equal seeds give bit-equal initial weights.  Each module owns its Adam in
``.optim`` like the reference.  ``net_spec()`` describes the module to the CUDA kernels (include/pspde.h); the
``forward`` methods are plain PyTorch and are used for host-side evaluation, never inside the training rollout.
"""
import torch as pt

from . import _lib as L


class SingleParam(pt.nn.Module):
    """Learnable scalar Y_0 (learn_Y_0=True)."""

    def __init__(self, lr, initial=None, seed=42):
        super().__init__()
        pt.manual_seed(seed)
        if initial is None:
            value = pt.tensor([0.0])
        elif initial == "random":
            value = pt.randn(1)
        else:
            value = pt.tensor([float(initial)])
        self.Y_0 = pt.nn.Parameter(value, requires_grad=True)
        self.register_parameter("param", self.Y_0)
        self.optim = pt.optim.Adam(self.parameters(), lr=lr)

    def forward(self, x):
        return self.Y_0


class DenseNet(pt.nn.Module):
    """Densely connected MLP: layer i sees the concatenation of the input and all earlier hidden outputs;
    hidden activation relu(.)**2; weights (fan_in_total, fan_out) ~ 0.1 N(0,1), zero biases."""

    def __init__(self, d_in, d_out, lr, arch=(30, 30), seed=42):
        super().__init__()
        pt.manual_seed(seed)
        self.nn_dims = [d_in] + list(arch) + [d_out]
        self.W = []
        for i in range(len(self.nn_dims) - 1):
            fan_in = sum(self.nn_dims[:i + 1])
            self.W.append(pt.nn.Parameter(pt.randn(fan_in, self.nn_dims[i + 1]) * 0.1))
            self.W.append(pt.nn.Parameter(pt.zeros(self.nn_dims[i + 1])))
        for i, w in enumerate(self.W):
            self.register_parameter("param %d" % i, w)
        self.optim = pt.optim.Adam(self.parameters(), lr=lr)

    def forward(self, x):
        n = len(self.nn_dims) - 1
        for i in range(n):
            pre = x @ self.W[2 * i] + self.W[2 * i + 1]
            if i == n - 1:
                return pre
            x = pt.cat([x, pt.relu(pre) ** 2], dim=1)
        return x

    def net_spec(self):
        return L.NET_DENSENET, list(self.nn_dims)


class MySequential(pt.nn.Module):
    """Plain MLP [d_in, 30, 30, d_out] with tanh, all weights and biases ~ N(0, 0.01^2)."""

    def __init__(self, d_in, d_out, lr, seed):
        super().__init__()
        pt.manual_seed(seed)
        self.nn_dims = [d_in, 30, 30, d_out]
        self.linears = pt.nn.ModuleList(
            [pt.nn.Linear(self.nn_dims[i], self.nn_dims[i + 1]) for i in range(len(self.nn_dims) - 1)])
        self.activations = pt.nn.ModuleList([pt.nn.Tanh() for _ in range(len(self.nn_dims) - 2)])
        self.optim = pt.optim.Adam(self.parameters(), lr=lr)
        for lin in self.linears:
            pt.nn.init.normal_(lin.weight, 0, 0.01)
            pt.nn.init.normal_(lin.bias, 0, 0.01)

    def forward(self, x):
        last = len(self.linears) - 1
        for i, lin in enumerate(self.linears):
            x = lin(x)
            if i < last:
                x = self.activations[i](x)
        return x

    def net_spec(self):
        return L.NET_MLP_TANH, list(self.nn_dims)
