"""models.lib.quantizer.VectorQuantizer on the B200 kernels (reference: code/models/lib/quantizer.py:14-90).

Same constructor, parameter name (`embedding.weight`, init U(+-1/n_e)) and return tuples.  The nearest-neighbour search
(dim_vq_argmin) and the row lookup (dim_vq_gather) run in libdimb200; the loss / perplexity bookkeeping that the
reference returns alongside stays as a few torch element-wise ops (reporting only)."""
import torch
import torch.nn as nn

from dim_b200 import ops


class VectorQuantizer(nn.Module):
    def __init__(self, n_e, e_dim, beta):
        super().__init__()
        self.n_e, self.e_dim, self.beta = n_e, e_dim, beta
        self.embedding = nn.Embedding(n_e, e_dim)
        self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)

    def _codebook(self):
        return self.embedding.weight.detach().contiguous()

    def forward(self, z):
        """z (B,L,D) -> (z_q (B,D,L), loss, (perplexity, min_encodings (N,K), min_encoding_indices (N,1)))."""
        zf = z.detach().reshape(-1, self.e_dim).contiguous()
        E = self._codebook()
        idx = ops.vq_argmin(zf, E)
        rows = ops.vq_gather(idx, E).view(z.shape)
        return self._package(z, rows, idx)

    def _package(self, z, rows, idx):
        min_encodings = torch.zeros(idx.shape[0], self.n_e, device=z.device, dtype=z.dtype)
        min_encodings.scatter_(1, idx.unsqueeze(1), 1)
        loss = self.beta * torch.mean((rows - z) ** 2) + torch.mean((rows - z) ** 2)
        z_q = z + (rows - z)                                   # the reference's straight-through value (quantizer.py:58)
        e_mean = torch.mean(min_encodings, dim=0)
        perplexity = torch.exp(-torch.sum(e_mean * torch.log(e_mean + 1e-10)))
        return z_q.permute(0, 2, 1).contiguous(), loss, (perplexity, min_encodings, idx.unsqueeze(1))

    def get_distance(self, z):
        """z (B,D,L) -> d (B,K,L): full distance table (diagnostic API; plain torch, not on the hot path)."""
        z = z.permute(0, 2, 1).contiguous()
        zf = z.view(-1, self.e_dim)
        E = self.embedding.weight
        d = torch.sum(zf ** 2, dim=1, keepdim=True) + torch.sum(E ** 2, dim=1) - 2 * torch.matmul(zf, E.t())
        return torch.reshape(d, (z.shape[0], -1, z.shape[2])).permute(0, 2, 1).contiguous()

    def get_codebook_entry(self, indices, shape):
        z_q = ops.vq_gather(indices.reshape(-1).long().contiguous(), self._codebook())
        return z_q.view(shape) if shape is not None else z_q
