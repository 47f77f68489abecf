"""models.stage1_BIWI.VQAutoEncoder on the B200 kernels (reference: code/models/stage1_BIWI.py:10-137, :254-393).

`encoder` / `decoder` are parameter containers with the reference's state_dict keys; encode/decode run through
dim_b200.engine.VQEngine (libdimb200: dim_vqvae_encode / dim_vqvae_decode).  Return tuples and layouts are the
reference's: encode -> (quant (B,C,L), emb_loss, (perplexity, one-hot (N,K), indices (N,1))), decode (B,C,L) -> (B,L,in_dim).
"""
import torch
import torch.nn.functional as F

from base import BaseModel
from dim_b200 import ops
from dim_b200.engine import PREC_FP32_TC, Handle, VQEngine
from dim_b200.paramtree import ParamTree, fingerprint, strip_prefix
from dim_b200.schema import SPEAKER_OUT_DIMS, VQConfig, vqspeaker_schema, vqvae_schema
from models.lib.quantizer import VectorQuantizer


class VQAutoEncoder(BaseModel):
    precision = PREC_FP32_TC          # fp32-grade arithmetic on the tensor cores (bit-exact code indices)

    def __init__(self, args):
        super().__init__()
        self.args = args
        self._cfg = VQConfig.from_cfg(args)
        schema = vqvae_schema(self._cfg)
        self.encoder = ParamTree(strip_prefix(schema, "encoder."))
        self.decoder = ParamTree(strip_prefix(schema, "decoder."))
        self.quantize = VectorQuantizer(args.n_embed, args.zquant_dim, beta=0.25)
        self._engine = None
        self._fp = None

    # ---- engine binding (weights are borrowed by the library: rebuild when storage changes) ----
    def engine(self) -> VQEngine:
        fp = fingerprint(self)
        if self._engine is None or fp != self._fp:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("VQAutoEncoder runs on CUDA only (sm_100a kernels, no CPU fallback): call .cuda() first")
            h = Handle(dev.index)
            h.register(self.state_dict())
            self._engine = VQEngine(h, self._cfg, precision=self.precision)
            self._fp = fp
        return self._engine

    # ---- reference API ----
    @torch.no_grad()
    def encode(self, x, x_a=None, lens=None, batch_index=None):
        idx, z, _ = self.engine().encode(x.float(), lens=lens, batch_index=batch_index, want_z=True)
        rows = ops.vq_gather(idx.reshape(-1), self.quantize.embedding.weight.detach().contiguous()).view(z.shape)
        return self.quantize._package(z, rows, idx.reshape(-1))

    @torch.no_grad()
    def decode(self, quant, batch_index=None):
        return self.engine().decode(quant=quant.float(), batch_index=batch_index)

    def forward(self, x):
        quant, emb_loss, info = self.encode(x)
        return self.decode(quant), emb_loss, info

    def sample_step(self, x, x_a=None):
        quant_z, _, info = self.encode(x, x_a)
        x_sample_det = self.decode(quant_z)
        btc = quant_z.shape[0], quant_z.shape[2], quant_z.shape[1]
        return x_sample_det, self.decode_to_img(info[2], btc)

    def get_quant(self, x, x_a=None):
        quant_z, _, info = self.encode(x, x_a)
        return quant_z, info[2]

    @torch.no_grad()
    def get_distances(self, x):
        _, z, _ = self.engine().encode(x.float(), want_z=True)
        return self.quantize.get_distance(z.permute(0, 2, 1))

    def get_quant_from_d(self, d, btc):
        return self.decode_to_img(torch.argmin(d, dim=1).unsqueeze(1), btc)

    @torch.no_grad()
    def entry_to_feature(self, index, zshape):
        return torch.reshape(self.quantize.get_codebook_entry(index.long().reshape(-1), shape=None), zshape)

    @torch.no_grad()
    def decode_to_img(self, index, zshape, batch_index=None):
        """codes -> frames; the gather is fused in front of the decoder (dim_vqvae_decode with codes)."""
        B, L = zshape[0], zshape[1]
        return self.engine().decode(codes=index.long().reshape(B, L), batch_index=batch_index)

    @torch.no_grad()
    def decode_logit(self, logits, zshape):
        ix = torch.topk(F.softmax(logits, dim=-1), k=1, dim=-1)[1] if logits.dim() == 3 else logits
        return self.decode_to_img(torch.reshape(ix, (-1, 1)), zshape)

    def get_logit(self, logits, sample=True, filter_value=-float("Inf"), temperature=0.7, top_p=0.9, sample_idx=None):
        probs = F.softmax(logits / temperature, dim=-1)
        if sample:
            s = probs.shape
            ix = torch.multinomial(probs.reshape(s[0] * s[1], s[2]), num_samples=1).reshape(s[0], s[1])
        else:
            ix = torch.topk(probs, k=1, dim=-1)[1]
        return ix, probs


class VQSpeakerAutoEncoder(VQAutoEncoder):
    """models.stage1_BIWI.VQSpeakerAutoEncoder (reference: code/models/stage1_BIWI.py:140-251; arch `stage1_BIWI_speaker`,
    code/config_speaker_old.yaml): one TransformerEncoder over the 824-d speaker frame (56 motion | 768 audio), 8 codes per
    frame from one 512 x 128 codebook, and TWO TransformerDecoders whose outputs are concatenated (decoder_v -> 56, decoder_a ->
    768).  On the B200 kernels this is three engines over the same handle (dim_vqvae_build_parts): the encoder half and one
    decoder-only model per output head; every other method is inherited (their bodies only call encode / decode)."""

    def __init__(self, args):
        BaseModel.__init__(self)
        self.args = args
        self._cfg = VQConfig.from_cfg(args)
        schema = vqspeaker_schema(self._cfg)
        self.encoder = ParamTree(strip_prefix(schema, "encoder."))
        self.decoder_v = ParamTree(strip_prefix(schema, "decoder_v."))
        self.decoder_a = ParamTree(strip_prefix(schema, "decoder_a."))
        self.quantize = VectorQuantizer(args.n_embed, args.zquant_dim, beta=0.25)
        self._engine = None
        self._fp = None

    def engine(self):
        fp = fingerprint(self)
        if self._engine is None or fp != self._fp:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("VQSpeakerAutoEncoder runs on CUDA only (sm_100a kernels, no CPU fallback): call .cuda() first")
            h = Handle(dev.index)
            h.register(self.state_dict())
            enc = VQEngine(h, self._cfg, precision=self.precision, encoder="encoder", decoder=None)
            decs = [VQEngine(h, self._cfg, precision=self.precision, encoder=None, decoder=name, out_dim=od)
                    for name, od in SPEAKER_OUT_DIMS]
            self._engine = (enc, decs)
            self._fp = fp
        return self._engine

    @torch.no_grad()
    def encode(self, x, x_a=None, lens=None, batch_index=None):
        idx, z, _ = self.engine()[0].encode(x.float(), lens=lens, batch_index=batch_index, want_z=True)
        z = z.view(z.shape[0], -1, self._cfg.zquant_dim)                    # h.view(B, -1, zquant_dim): 8 tokens per frame
        rows = ops.vq_gather(idx.reshape(-1), self.quantize.embedding.weight.detach().contiguous()).view(z.shape)
        return self.quantize._package(z, rows, idx.reshape(-1))

    @torch.no_grad()
    def decode(self, quant, batch_index=None):
        q = quant.float()
        return torch.cat([d.decode(quant=q, batch_index=batch_index) for d in self.engine()[1]], dim=-1)

    @torch.no_grad()
    def get_distances(self, x):
        _, z, _ = self.engine()[0].encode(x.float(), want_z=True)
        return self.quantize.get_distance(z)

    @torch.no_grad()
    def decode_to_img(self, index, zshape, batch_index=None):
        B, L = zshape[0], zshape[1]                                         # L counts codes: frames * face_quan_num
        codes = index.long().reshape(B, L)
        return torch.cat([d.decode(codes=codes, batch_index=batch_index) for d in self.engine()[1]], dim=-1)
