"""Op-by-op restatement of one forward pass.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Independent of ``nn.Transformer``:
every matmul, softmax and LayerNorm is written out, working from a plain
``state_dict``.  Follows models/transformer.py:47-68 for the wrapper and
torch.nn.Transformer (post-norm, ReLU, packed in-proj, eps 1e-5, final
encoder/decoder norms - constructed at models/transformer.py:38-44) for the layers.

``operand`` emulates the GEMM operand formats of the CUDA library so precision
choices can be studied on CPU:  "fp32" (exact), "fp16"/"bf16" (operands rounded,
fp32 accumulate), "fp16x2" (hi + 2^-11*lo split, three products, the ll term
dropped - what the library's split mode computes).
"""
import math

import torch

LN_EPS = 1e-5
SPLIT_SCALE = 2048.0  # 2^11


def _round_to(x, dt):
    return x.to(dt).to(x.dtype)


def gemm(x, w, b=None, operand="fp32"):
    """x (M,K) @ w (N,K)^T + b, with operand-format emulation."""
    if operand == "fp32":
        y = x @ w.t()
    elif operand in ("fp16", "bf16"):
        dt = torch.float16 if operand == "fp16" else torch.bfloat16
        y = _round_to(x, dt) @ _round_to(w, dt).t()
    elif operand == "fp16x2":
        xh = _round_to(x, torch.float16)
        xl = _round_to((x - xh) * SPLIT_SCALE, torch.float16)
        wh = _round_to(w, torch.float16)
        wl = _round_to((w - wh) * SPLIT_SCALE, torch.float16)
        y = xh @ wh.t() + (xh @ wl.t() + xl @ wh.t()) * (1.0 / SPLIT_SCALE)
    else:
        raise ValueError(operand)
    if b is not None:
        y = y + b
    return y


def layer_norm(x, w, b):
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)  # biased
    return (x - mean) / torch.sqrt(var + LN_EPS) * w + b


def _no_drop(site, x):
    return x


def mha(q_in, kv_in, sd, prefix, n_heads, mask, operand, drop=_no_drop, site_p=0, site_out=0):
    """q_in (B,Sq,d), kv_in (B,Sk,d); mask (Sq,Sk) additive or None.  drop(site, x2d): training-mode dropout hook
    (oracle/dropout.py) on the attention probabilities (rows (b, h, query), cols key) and on the projected output."""
    B, Sq, d = q_in.shape
    Sk = kv_in.shape[1]
    hd = d // n_heads
    W = sd[prefix + "in_proj_weight"]
    c = sd[prefix + "in_proj_bias"]
    q = gemm(q_in.reshape(-1, d), W[0:d], c[0:d], operand).view(B, Sq, n_heads, hd)
    k = gemm(kv_in.reshape(-1, d), W[d:2 * d], c[d:2 * d], operand).view(B, Sk, n_heads, hd)
    v = gemm(kv_in.reshape(-1, d), W[2 * d:], c[2 * d:], operand).view(B, Sk, n_heads, hd)
    s = torch.einsum("bqhe,bkhe->bhqk", q, k) / math.sqrt(hd)
    if mask is not None:
        s = s + mask
    p = torch.softmax(s, dim=-1)
    p = drop(site_p, p.reshape(B * n_heads * Sq, Sk)).view(B, n_heads, Sq, Sk)
    o = torch.einsum("bhqk,bkhe->bqhe", p, v).reshape(B * Sq, d)
    o = drop(site_out, gemm(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"], operand))
    return o.view(B, Sq, d)


def ffn(x, sd, prefix, operand, drop=_no_drop, site_h=0, site_out=0):
    B, S, d = x.shape
    h = torch.relu(gemm(x.reshape(-1, d), sd[prefix + "linear1.weight"], sd[prefix + "linear1.bias"], operand))
    h = drop(site_h, h)
    return drop(site_out, gemm(h, sd[prefix + "linear2.weight"], sd[prefix + "linear2.bias"], operand)).view(B, S, d)


def count_layers(sd, which):
    n = 0
    while f"transformer.{which}.layers.{n}.norm1.weight" in sd:
        n += 1
    return n


def embed(x, sd, pe_index, operand, drop=_no_drop, site=0):
    """emb(x)[b,s] = (x[b,s] W^T + bias) * sqrt(d) + PE[pe_index[b]]   (models/transformer.py:53-56), then the
    dropout of positional_encoding.py:35."""
    B, S, E = x.shape
    W = sd["embedding.weight"]
    d = W.shape[0]
    e = gemm(x.reshape(-1, E), W, sd["embedding.bias"], operand).view(B, S, d) * math.sqrt(d)
    pe = sd["positional_encoder.pos_encoding"][:, 0, :].to(x.dtype)
    return drop(site, (e + pe[pe_index][:, None, :]).reshape(B * S, d)).view(B, S, d)


def forward(sd, src, tgt, n_heads, tgt_mask=None, pe_index=None, operand="fp32", hp_first=False, drop=None):
    """Returns (S_tgt, B, E) like models/transformer.py:65-68.

    ``pe_index`` (B,) long: PE row per clip; default arange(B) = the reference.
    ``hp_first``: keep the embedding and layer-0 Q/K/V projections in fp16x2 even
    when ``operand`` is a 16-bit format (the library's "mixed" mode).
    ``drop``: None (eval mode) or an oracle.dropout.Dropper - training-mode dropout at nn.Transformer's sites.
    """
    from . import dropout as D
    if drop is None:
        drop = _no_drop
    elif hp_first and operand in ("fp16", "bf16"):
        raise ValueError("dropout hooks are restated for the plain operand path only")
    B = src.shape[0]
    if pe_index is None:
        if B > 64:
            raise RuntimeError("reference semantics: B <= 64 (PositionalEncoding max_len, models/transformer.py:33-35)")
        pe_index = torch.arange(B)
    Le, Ld = count_layers(sd, "encoder"), count_layers(sd, "decoder")
    op0 = "fp16x2" if (hp_first and operand in ("fp16", "bf16")) else operand

    def mha_l(q_in, kv_in, prefix, mask, first, site_p=0, site_out=0):
        if not first or op0 == operand:
            return mha(q_in, kv_in, sd, prefix, n_heads, mask, operand, drop, site_p, site_out)
        # layer-0 self-attention: QKV in high precision, out-proj in `operand`
        Bq, Sq, d = q_in.shape
        hd = d // n_heads
        W, c = sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"]
        qkv = gemm(q_in.reshape(-1, d), W, c, op0).view(Bq, Sq, 3, n_heads, hd)
        s = torch.einsum("bqhe,bkhe->bhqk", qkv[:, :, 0], qkv[:, :, 1]) / math.sqrt(hd)
        if mask is not None:
            s = s + mask
        o = torch.einsum("bhqk,bkhe->bqhe", torch.softmax(s, -1), qkv[:, :, 2]).reshape(Bq * Sq, d)
        return gemm(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"], operand).view(Bq, Sq, d)

    x = embed(src, sd, pe_index, op0, drop, 0)
    for l in range(Le):
        p = f"transformer.encoder.layers.{l}."
        x = layer_norm(x + mha_l(x, x, p + "self_attn.", None, l == 0, D.enc_site(l, D.SA_P), D.enc_site(l, D.SA_OUT)),
                       sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        x = layer_norm(x + ffn(x, sd, p, operand, drop, D.enc_site(l, D.FF_H), D.enc_site(l, D.FF_OUT)),
                       sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    mem = layer_norm(x, sd["transformer.encoder.norm.weight"], sd["transformer.encoder.norm.bias"])

    y = embed(tgt, sd, pe_index, op0, drop, 1)
    for l in range(Ld):
        p = f"transformer.decoder.layers.{l}."
        y = layer_norm(y + mha_l(y, y, p + "self_attn.", tgt_mask, l == 0, D.dec_site(l, D.SA_P), D.dec_site(l, D.SA_OUT)),
                       sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        y = layer_norm(y + mha(y, mem, sd, p + "multihead_attn.", n_heads, None, operand, drop, D.dec_site(l, D.CA_P),
                               D.dec_site(l, D.CA_OUT)),
                       sd[p + "norm2.weight"], sd[p + "norm2.bias"])
        y = layer_norm(y + ffn(y, sd, p, operand, drop, D.dec_site(l, D.FF_H), D.dec_site(l, D.FF_OUT)),
                       sd[p + "norm3.weight"], sd[p + "norm3.bias"])
    y = layer_norm(y, sd["transformer.decoder.norm.weight"], sd["transformer.decoder.norm.bias"])
    Bt, St, d = y.shape
    out = gemm(y.reshape(-1, d), sd["out.weight"], sd["out.bias"], operand).view(Bt, St, -1)
    return out.permute(1, 0, 2).contiguous()


def causal_mask(size, dtype=torch.float32):
    """models/transformer.py:70-89 - 0 on/below the diagonal, -inf above."""
    m = torch.zeros(size, size, dtype=dtype)
    return m.masked_fill(torch.triu(torch.ones(size, size, dtype=torch.bool), 1), float("-inf"))
