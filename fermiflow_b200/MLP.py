"""Radial MLP of the backflow potentials -- host-side mirror of reference src/MLP.py:4-45.

Same parameter names and shapes as the reference (fc1: Linear(D_in, D_hidden), fc2:
Linear(D_hidden, 1, bias=False), sigmoid), so reference checkpoints load unchanged.  The
module itself is only a parameter container: on the hot path the CUDA kernels read
fc1.weight / fc1.bias / fc2.weight directly (see Backflow, CNF).  forward()/grad() are kept
as small torch expressions for inspection and plotting (flow.py:86 backflow_potential).
"""
import torch


class MLP(torch.nn.Module):
    def __init__(self, D_in, D_hidden):
        super().__init__()
        self.fc1 = torch.nn.Linear(D_in, D_hidden, dtype=torch.float64)
        self.fc2 = torch.nn.Linear(D_hidden, 1, bias=False, dtype=torch.float64)
        self.activation = torch.nn.Sigmoid()

    def init_zeros(self):                                   # MLP.py:19-22
        for p in (self.fc1.weight, self.fc1.bias, self.fc2.weight):
            torch.nn.init.zeros_(p)

    def init_gaussian(self, seed, std=1e-3):                # MLP.py:24-29
        torch.manual_seed(seed)
        for p in (self.fc1.weight, self.fc1.bias, self.fc2.weight):
            torch.nn.init.normal_(p, std=std)

    def forward(self, x):                                   # MLP.py:31-33
        return self.fc2(self.activation(self.fc1(x)))

    def d_sigmoid(self, output):
        return output * (1. - output)

    def grad(self, x):                                      # MLP.py:38-45
        s = self.activation(self.fc1(x))
        return (self.fc2.weight * self.d_sigmoid(s)).matmul(self.fc1.weight)

    def kernel_params(self):
        """(w1, b1, w2) float64 vectors as the kernels want them (D_in must be 1)."""
        if self.fc1.in_features != 1:
            raise ValueError("the CUDA path evaluates radial MLPs: D_in must be 1, got %d" % self.fc1.in_features)
        return (self.fc1.weight.detach().reshape(-1).contiguous(), self.fc1.bias.detach().contiguous(),
                self.fc2.weight.detach().reshape(-1).contiguous())

    def parameters_in_kernel_order(self):
        return (self.fc1.weight, self.fc1.bias, self.fc2.weight)
