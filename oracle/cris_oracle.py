"""CPU fp32 restatement ("oracle") of the CRIS hot path — TEST INFRASTRUCTURE ONLY.

This file restates, in plain functional PyTorch fp32 on CPU, the algorithm of the reference's
`model.segmenter.CRIS.forward` (DerrickWang005/CRIS.pytorch) so that the hand-written sm_100a
kernels of cris.pytorch_b200 have something to be checked against on a box where
/root/reference does not exist.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it; the product path never does.

Pinning: the reference ships NO tests, golden vectors or fixtures for this path (SURVEY.md §4,
§8c).  The oracle is therefore pinned against outputs of the reference itself, generated in the
build container by oracle/make_golden.py (which imports /root/reference/model unchanged) and
committed under tests/golden/; tests/test_oracle_golden.py replays them.

Every function cites the reference file:line it restates.  Token tensors are batch-first
[B, L, C] here (the reference is sequence-first); weights use the reference's state_dict names.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

BN_EPS = 1e-5
LN_EPS = 1e-5


class OracleState:
    """Carries mode flags and collects updated BatchNorm running statistics."""

    def __init__(self, sd: SD, training: bool, dropout_p: float = 0.0, taps: Optional[dict] = None,
                 storage: str = "fp32"):
        self.sd = sd
        self.training = training
        self.dropout_p = dropout_p
        self.new_running: Dict[str, Tensor] = {}
        self.taps = taps  # optional dict collecting named intermediates
        # storage = "fp32": the reference algorithm in full precision (what the golden fixtures pin).
        # storage = "bf16": the SAME algorithm with every tensor that cris.pytorch_b200 keeps in bf16
        # (GEMM operands/outputs, activations) rounded to bf16 where it is stored — the
        # same-precision comparison that autocast gives the reference (engine/engine.py:48).  Rounding uses
        # a straight-through gradient so autograd still yields the reference gradients at those activations.
        assert storage in ("fp32", "bf16")
        self.storage = storage

    def q(self, x: Tensor) -> Tensor:
        """activation storage rounding"""
        if self.storage == "fp32":
            return x
        return x + (x.to(torch.bfloat16).float() - x).detach()

    def w(self, name: str) -> Tensor:
        """GEMM weight operand (conv / linear / in_proj / text_projection are bf16 tensor-core operands)"""
        return self.q(self.sd[name])

    def tap(self, name: str, t: Tensor) -> Tensor:
        if self.taps is not None:
            self.taps[name] = t
        return t


# --------------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------------
def batch_norm(st: OracleState, x: Tensor, prefix: str) -> Tensor:
    """nn.BatchNorm2d / BatchNorm1d (model/clip.py:18,21,26,171-183; model/layers.py:8-16,262).

    Training mode = batch statistics (biased variance for the normalisation, unbiased for the
    running estimate, momentum 0.1) — engine/engine.py:29 puts every BN, incl. the CLIP trunk,
    in this mode.  Eval mode = running statistics.
    """
    sd = st.sd
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    dims = [0] + list(range(2, x.dim()))
    shape = [1, -1] + [1] * (x.dim() - 2)
    if st.training:
        mean = x.mean(dim=dims)
        var = x.var(dim=dims, unbiased=False)
        n = x.numel() // x.shape[1]
        with torch.no_grad():
            rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
            st.new_running[prefix + ".running_mean"] = 0.9 * rm + 0.1 * mean.detach()
            st.new_running[prefix + ".running_var"] = 0.9 * rv + 0.1 * var.detach() * (n / max(n - 1, 1))
    else:
        mean, var = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    inv = torch.rsqrt(var + BN_EPS)
    return (x - mean.view(shape)) * (inv * w).view(shape) + b.view(shape)


def layer_norm(st: OracleState, x: Tensor, prefix: str) -> Tensor:
    """nn.LayerNorm over the last dim, eps 1e-5 (model/clip.py:226-231; model/layers.py:199-216)."""
    w, b = st.sd[prefix + ".weight"], st.sd[prefix + ".bias"]
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + LN_EPS) * w + b


def dropout(st: OracleState, x: Tensor) -> Tensor:
    """nn.Dropout (model/layers.py:202-219).  Parity runs use p = 0."""
    if st.training and st.dropout_p > 0:
        return F.dropout(x, st.dropout_p, True)
    return x


def attention(q: Tensor, k: Tensor, v: Tensor, heads: int, add_mask: Optional[Tensor] = None,
              key_padding: Optional[Tensor] = None, st: Optional[OracleState] = None) -> Tensor:
    """Scaled dot-product attention core of F.multi_head_attention_forward as the reference calls
    it (model/clip.py:119-139,255-260; model/layers.py:235,240-243): per-head softmax(q k^T /
    sqrt(hd) + additive mask, key padding -> -inf), dropout on the probabilities, times v.
    q: [B, Lq, E]; k, v: [B, Lk, E]  ->  [B, Lq, E]
    """
    B, Lq, E = q.shape
    Lk = k.shape[1]
    hd = E // heads
    qh = q.view(B, Lq, heads, hd).transpose(1, 2)
    kh = k.view(B, Lk, heads, hd).transpose(1, 2)
    vh = v.view(B, Lk, heads, hd).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) * (1.0 / math.sqrt(hd))
    if st is not None:
        s = st.q(s)
    if add_mask is not None:
        s = s + add_mask
    if key_padding is not None:
        s = s.masked_fill(key_padding[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    if st is not None:
        p = st.q(dropout(st, st.q(p)))
    o = (p @ vh).transpose(1, 2).reshape(B, Lq, E)
    return st.q(o) if st is not None else o


# --------------------------------------------------------------------------------------------
# image encoder — model/clip.py:10-57 (Bottleneck), :60-144 (AttentionPool2d), :147-223
# --------------------------------------------------------------------------------------------
def bottleneck(st: OracleState, x: Tensor, p: str, stride: int) -> Tensor:
    """model/clip.py:44-57: 1x1 -> BN -> ReLU -> 3x3 (stride 1) -> BN -> ReLU -> AvgPool(stride)
    -> 1x1 -> BN, identity through [AvgPool(stride) -> 1x1 -> BN] when shapes change, add, ReLU."""
    sd, q = st.sd, st.q
    out = q(F.relu(batch_norm(st, q(F.conv2d(x, st.w(p + ".conv1.weight"))), p + ".bn1")))
    out = q(F.relu(batch_norm(st, q(F.conv2d(out, st.w(p + ".conv2.weight"), padding=1)), p + ".bn2")))
    if stride > 1:
        out = q(F.avg_pool2d(out, stride))
    out = batch_norm(st, q(F.conv2d(out, st.w(p + ".conv3.weight"))), p + ".bn3")
    if (p + ".downsample.0.weight") in sd:
        idn = q(F.avg_pool2d(x, stride)) if stride > 1 else x
        idn = q(batch_norm(st, q(F.conv2d(idn, st.w(p + ".downsample.0.weight"))), p + ".downsample.1"))
    else:
        idn = x
    return q(F.relu(out + idn))


def resized_pos_embed(pos: Tensor, spacial: int, hw) -> Tensor:
    """model/clip.py:80-108: drop the CLS row, view [1,C,s,s], bicubic (align_corners=False) to
    (H,W), return [H*W, C]."""
    C = pos.shape[1]
    grid = pos[1:].reshape(1, spacial, spacial, C).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, size=hw, mode="bicubic", align_corners=False)
    return grid.flatten(2)[0].t()


def attention_pool(st: OracleState, x: Tensor, p: str, heads: int) -> Tensor:
    """model/clip.py:110-144: residual conv1x1+BN, add resized pos-embed, MHA over all H*W tokens
    (separate q/k/v weights, no dropout), c_proj, + residual, ReLU."""
    sd = st.sd
    B, C, H, W = x.shape
    rq = st.q
    res = batch_norm(st, rq(F.conv2d(x, st.w(p + ".connect.0.weight"))), p + ".connect.1")
    spacial = int(round(math.sqrt(sd[p + ".positional_embedding"].shape[0] - 1)))
    tok = rq(x.flatten(2).transpose(1, 2) + resized_pos_embed(sd[p + ".positional_embedding"], spacial, (H, W)))
    q = rq(F.linear(tok, st.w(p + ".q_proj.weight"), sd[p + ".q_proj.bias"]))
    k = rq(F.linear(tok, st.w(p + ".k_proj.weight"), sd[p + ".k_proj.bias"]))
    v = rq(F.linear(tok, st.w(p + ".v_proj.weight"), sd[p + ".v_proj.bias"]))
    o = attention(q, k, v, heads, st=st if st.storage != "fp32" else None)
    o = rq(F.linear(o, st.w(p + ".c_proj.weight"), sd[p + ".c_proj.bias"]))
    o = o.transpose(1, 2).reshape(B, -1, H, W)
    return rq(F.relu(o + res))


def encode_image(st: OracleState, img: Tensor):
    """model/clip.py:207-223 (+ :436-437): 3-conv stem, avgpool, layer1..4, attnpool -> (C3,C4,C5)."""
    sd = st.sd
    v = "backbone.visual"
    q = st.q
    # stem conv1 = bf16 im2col patches x bf16 weights on the tensor cores (fp32 accumulate), like every conv
    x = q(F.relu(batch_norm(st, q(F.conv2d(q(img), st.w(v + ".conv1.weight"), stride=2, padding=1)), v + ".bn1")))
    x = q(F.relu(batch_norm(st, q(F.conv2d(x, st.w(v + ".conv2.weight"), padding=1)), v + ".bn2")))
    x = q(F.relu(batch_norm(st, q(F.conv2d(x, st.w(v + ".conv3.weight"), padding=1)), v + ".bn3")))
    x = q(F.avg_pool2d(x, 2))
    st.tap("stem", x)
    feats = []
    for li in (1, 2, 3, 4):
        bi = 0
        while f"{v}.layer{li}.{bi}.conv1.weight" in sd:
            stride = 2 if (li > 1 and bi == 0) else 1
            x = bottleneck(st, x, f"{v}.layer{li}.{bi}", stride)
            bi += 1
        st.tap(f"layer{li}", x)
        feats.append(x)
    width = sd[v + ".layer1.0.conv1.weight"].shape[0]
    heads = width * 32 // 64  # model/clip.py:356
    c5 = attention_pool(st, feats[3], v + ".attnpool", heads)
    st.tap("attnpool", c5)
    return feats[1], feats[2], c5


# --------------------------------------------------------------------------------------------
# text encoder — model/clip.py:239-283, :424-456
# --------------------------------------------------------------------------------------------
def encode_text(st: OracleState, word: Tensor):
    """model/clip.py:439-456: token + positional embedding, pre-LN transformer blocks with a
    causal -inf mask (:424-430), QuickGELU MLP (:234-236), ln_final; state = feature at the EOT
    position (argmax of the token ids) times text_projection."""
    sd = st.sd
    b = "backbone"
    B, L = word.shape
    x = sd[b + ".token_embedding.weight"][word] + sd[b + ".positional_embedding"][:L]
    causal = torch.full((L, L), float("-inf")).triu_(1)
    width = x.shape[-1]
    heads = width // 64  # model/clip.py:538
    i = 0
    while f"{b}.transformer.resblocks.{i}.ln_1.weight" in sd:
        p = f"{b}.transformer.resblocks.{i}"
        rq = st.q
        h = rq(layer_norm(st, x, p + ".ln_1"))
        qkv = rq(F.linear(h, st.w(p + ".attn.in_proj_weight"), sd[p + ".attn.in_proj_bias"]))
        q, k, v = qkv.split(width, dim=-1)
        a = attention(q, k, v, heads, add_mask=causal, st=st if st.storage != "fp32" else None)
        x = x + rq(F.linear(a, st.w(p + ".attn.out_proj.weight"), sd[p + ".attn.out_proj.bias"]))
        h = rq(layer_norm(st, x, p + ".ln_2"))
        h = rq(F.linear(h, st.w(p + ".mlp.c_fc.weight"), sd[p + ".mlp.c_fc.bias"]))
        h = rq(h * torch.sigmoid(1.702 * h))
        x = x + rq(F.linear(h, st.w(p + ".mlp.c_proj.weight"), sd[p + ".mlp.c_proj.bias"]))
        i += 1
    x = st.q(layer_norm(st, x, b + ".ln_final"))
    eot = word.argmax(dim=-1)
    state = st.q(x[torch.arange(B), eot] @ st.w(b + ".text_projection"))
    st.tap("word", x)
    st.tap("state", state)
    return x, state


# --------------------------------------------------------------------------------------------
# FPN neck — model/layers.py:253-309
# --------------------------------------------------------------------------------------------
def conv_bn_relu(st: OracleState, x: Tensor, p: str, pad: int) -> Tensor:
    """model/layers.py:8-11 conv_layer: Conv2d(no bias) + BN + ReLU."""
    return st.q(F.relu(batch_norm(st, st.q(F.conv2d(x, st.w(p + ".0.weight"), padding=pad)), p + ".1")))


def up2(x: Tensor, st: Optional[OracleState] = None) -> Tensor:
    """bilinear x2, align_corners=False (model/layers.py:54,56,293,304)."""
    y = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    return st.q(y) if st is not None else y


def fpn(st: OracleState, c3: Tensor, c4: Tensor, c5: Tensor, state: Tensor) -> Tensor:
    """model/layers.py:282-309."""
    sd = st.sd
    q = st.q
    s = q(F.linear(state, st.w("neck.txt_proj.0.weight")))
    s = q(F.relu(batch_norm(st, s, "neck.txt_proj.1")))[:, :, None, None]
    f5 = conv_bn_relu(st, c5, "neck.f1_v_proj", 0)
    f5 = q(F.relu(batch_norm(st, q(f5 * s), "neck.norm_layer.0")))
    f4 = conv_bn_relu(st, c4, "neck.f2_v_proj", 1)
    f4 = conv_bn_relu(st, torch.cat([f4, up2(f5, st)], 1), "neck.f2_cat", 0)
    f3 = conv_bn_relu(st, c3, "neck.f3_v_proj", 1)
    f3 = q(F.avg_pool2d(f3, 2, 2))
    f3 = conv_bn_relu(st, torch.cat([f3, f4], 1), "neck.f3_cat", 0)
    fq5 = up2(conv_bn_relu(st, f5, "neck.f4_proj5", 1), st)
    fq4 = conv_bn_relu(st, f4, "neck.f4_proj4", 1)
    fq3 = conv_bn_relu(st, f3, "neck.f4_proj3", 1)
    fq = conv_bn_relu(st, torch.cat([fq3, fq4, fq5], 1), "neck.aggr", 0)
    # CoordConv (model/layers.py:30-44): append x then y in [-1,1]
    B, _, H, W = fq.shape
    ys = torch.linspace(-1, 1, H).view(1, 1, H, 1).expand(B, 1, H, W)
    xs = torch.linspace(-1, 1, W).view(1, 1, 1, W).expand(B, 1, H, W)
    fq = conv_bn_relu(st, torch.cat([fq, q(xs), q(ys)], 1), "neck.coordconv.0.conv1", 1)
    fq = conv_bn_relu(st, fq, "neck.coordconv.1", 1)
    return st.tap("fq", fq)


# --------------------------------------------------------------------------------------------
# vision-language decoder — model/layers.py:87-250
# --------------------------------------------------------------------------------------------
def sine_pos_1d(d: int, length: int) -> Tensor:
    """model/layers.py:106-123 -> [length, d]."""
    pos = torch.arange(length, dtype=torch.float32)[:, None]
    freq = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    pe = torch.zeros(length, d)
    pe[:, 0::2] = torch.sin(pos * freq)
    pe[:, 1::2] = torch.cos(pos * freq)
    return pe


def sine_pos_2d(d: int, H: int, W: int) -> Tensor:
    """model/layers.py:125-152 -> [H*W, d]: channels [0,d/2) encode w, [d/2,d) encode h; token = h*W+w."""
    half = d // 2
    freq = torch.exp(torch.arange(0.0, half, 2) * -(math.log(10000.0) / half))
    pw = torch.arange(0.0, W)[:, None] * freq  # [W, half/2]
    ph = torch.arange(0.0, H)[:, None] * freq
    pe = torch.zeros(H, W, d)
    pe[:, :, 0:half:2] = torch.sin(pw)[None, :, :]
    pe[:, :, 1:half:2] = torch.cos(pw)[None, :, :]
    pe[:, :, half::2] = torch.sin(ph)[:, None, :]
    pe[:, :, half + 1::2] = torch.cos(ph)[:, None, :]
    return pe.reshape(H * W, d)


def mha_proj(st: OracleState, p: str, q_in: Tensor, k_in: Tensor, v_in: Tensor):
    """Packed in_proj of nn.MultiheadAttention applied to three different inputs."""
    w, b = st.w(p + ".in_proj_weight"), st.sd[p + ".in_proj_bias"]
    E = w.shape[1]
    q = st.q(F.linear(q_in, w[:E], b[:E]))
    k = st.q(F.linear(k_in, w[E:2 * E], b[E:2 * E]))
    v = st.q(F.linear(v_in, w[2 * E:], b[2 * E:]))
    return q, k, v


def decoder(st: OracleState, fq: Tensor, word: Tensor, pad_mask: Tensor, heads: int) -> Tensor:
    """model/layers.py:154-188 (+ layer :224-250)."""
    sd = st.sd
    B, C, H, W = fq.shape
    L = word.shape[1]
    vpos = sine_pos_2d(C, H, W)
    tpos = sine_pos_1d(word.shape[2], L)
    rq = st.q
    vis = fq.flatten(2).transpose(1, 2)  # [B, HW, C]
    tk = rq(word + tpos)
    i = 0
    while f"decoder.layers.{i}.norm1.weight" in sd:
        p = f"decoder.layers.{i}"
        ln = layer_norm(st, vis, p + ".norm1")
        v2, v2p = rq(ln), rq(ln + vpos)
        q, k, v = mha_proj(st, p + ".self_attn", v2p, v2p, v2)
        a = attention(q, k, v, heads, st=st)
        a = rq(F.linear(a, st.w(p + ".self_attn.out_proj.weight"), sd[p + ".self_attn.out_proj.bias"]))
        vis = vis + dropout(st, rq(layer_norm(st, a, p + ".self_attn_norm")))
        v2p = rq(layer_norm(st, vis, p + ".norm2") + vpos)
        q, k, v = mha_proj(st, p + ".multihead_attn", v2p, tk, word)
        a = attention(q, k, v, heads, key_padding=pad_mask, st=st)
        a = rq(F.linear(a, st.w(p + ".multihead_attn.out_proj.weight"), sd[p + ".multihead_attn.out_proj.bias"]))
        vis = vis + dropout(st, rq(layer_norm(st, a, p + ".cross_attn_norm")))
        v2 = rq(layer_norm(st, vis, p + ".norm3"))
        h = rq(F.relu(F.linear(v2, st.w(p + ".ffn.0.weight"), sd[p + ".ffn.0.bias"])))
        h = rq(layer_norm(st, dropout(st, h), p + ".ffn.3"))
        h = rq(F.linear(h, st.w(p + ".ffn.4.weight"), sd[p + ".ffn.4.bias"]))
        vis = vis + dropout(st, h)
        st.tap(f"dec{i}", vis)
        i += 1
    vis = rq(layer_norm(st, vis, "decoder.norm"))
    return vis.transpose(1, 2).reshape(B, C, H, W)


# --------------------------------------------------------------------------------------------
# projector + loss — model/layers.py:47-84, model/segmenter.py:52-62
# --------------------------------------------------------------------------------------------
def projector(st: OracleState, fq: Tensor, state: Tensor) -> Tensor:
    """model/layers.py:63-84: up x2, conv3x3+BN+ReLU, up x2, conv3x3+BN+ReLU, conv1x1(+bias); the
    text Linear yields per-sample 3x3 kernels [B,C,3,3] and a bias [B]; per-sample correlation."""
    sd = st.sd
    x = conv_bn_relu(st, up2(fq, st), "proj.vis.1", 1)
    x = conv_bn_relu(st, up2(x, st), "proj.vis.3", 1)
    x = st.q(F.conv2d(x, st.w("proj.vis.4.weight"), sd["proj.vis.4.bias"]))
    st.tap("proj_feat", x)
    B, C, H, W = x.shape
    t = F.linear(state, st.w("proj.txt.weight"), sd["proj.txt.bias"])  # kept in fp32 by the engine too
    kern, bias = t[:, :-1].reshape(B, C, 3, 3), t[:, -1]
    patches = F.unfold(x, 3, padding=1).view(B, C * 9, H * W)
    out = torch.einsum("bkp,bk->bp", patches, kern.reshape(B, C * 9)) + bias[:, None]
    return out.view(B, 1, H, W)


def cris_forward(sd: SD, img: Tensor, word: Tensor, mask: Optional[Tensor] = None, *, training: bool = False,
                 num_head: int = 8, dropout_p: float = 0.0, taps: Optional[dict] = None, storage: str = "fp32"):
    """model/segmenter.py:29-62.  Returns dict(pred, mask, loss, new_running).
    storage="bf16" = same algorithm with the engine's bf16 storage points emulated (see OracleState)."""
    st = OracleState(sd, training, dropout_p, taps, storage)
    pad_mask = word == 0
    c3, c4, c5 = encode_image(st, img)
    wfeat, state = encode_text(st, word)
    fq = fpn(st, c3, c4, c5, state)
    fq = decoder(st, fq, wfeat, pad_mask, num_head)
    st.tap("dec_out", fq)
    pred = projector(st, fq, state)
    out = {"pred": pred, "new_running": st.new_running}
    if mask is not None:
        if pred.shape[-2:] != mask.shape[-2:]:
            # F.interpolate(..., 'nearest') == pixel (floor(i*scale), floor(j*scale))
            sh, sw = mask.shape[-2] // pred.shape[-2], mask.shape[-1] // pred.shape[-1]
            mask = mask[:, :, ::sh, ::sw]
        x, t = pred, mask
        loss = (torch.clamp(x, min=0) - x * t + torch.log1p(torch.exp(-x.abs()))).mean()
        out["mask"] = mask
        out["loss"] = loss
    return out
