"""Restatement of the reference model on top of ``torch.nn.Transformer``.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows, line by line in behaviour:

  * models/positional_encoding.py:7-35  -> ``RefPositionalEncoding``
  * models/transformer.py:12-45         -> ``RefTransformer.__init__``
  * models/transformer.py:47-68         -> ``RefTransformer.forward``
  * models/transformer.py:70-89         -> ``RefTransformer.get_tgt_mask``

The only deliberate difference is the constructor: the reference parses
``sys.argv`` and a yaml file to obtain ``FRAME_SIZE`` (models/transformer.py:23,
28-29); here it is the keyword ``frame_size``.  Sub-module names, creation order
(hence RNG consumption and random-init values) and ``state_dict`` keys are the
reference's.
"""
import math

import torch
import torch.nn as nn


class RefPositionalEncoding(nn.Module):
    def __init__(self, dim_model, dropout_p, max_len):
        super().__init__()
        self.dropout = nn.Dropout(dropout_p)
        # models/positional_encoding.py:17-25
        table = torch.zeros(max_len, dim_model)
        pos = torch.arange(0, max_len, dtype=torch.float).view(-1, 1)
        div = torch.exp(torch.arange(0, dim_model, 2).float() * (-math.log(10000.0)) / dim_model)
        table[:, 0::2] = torch.sin(pos * div)
        table[:, 1::2] = torch.cos(pos * div)
        # models/positional_encoding.py:28,31 -> (max_len, 1, d)
        self.register_buffer("pos_encoding", table.unsqueeze(0).transpose(0, 1))

    def forward(self, token_embedding):
        # models/positional_encoding.py:35 - sliced by dim 0 of a BATCH-FIRST tensor:
        # clip b receives pos_encoding[b] on every token.
        return self.dropout(token_embedding + self.pos_encoding[: token_embedding.size(0), :])


class RefTransformer(nn.Module):
    def __init__(self, num_tokens=0, dim_model=256, num_heads=8, num_encoder_layers=6,
                 num_decoder_layers=6, dropout_p=0.1, *, frame_size=64):
        super().__init__()
        self.dim_model = dim_model
        self.height = frame_size
        self.width = frame_size
        self.compression = 8
        latent = self.height // self.compression * self.width // self.compression * 4
        self.positional_encoder = RefPositionalEncoding(dim_model=dim_model, dropout_p=dropout_p, max_len=64)
        self.embedding = nn.Linear(latent, dim_model)
        self.transformer = nn.Transformer(d_model=dim_model, nhead=num_heads,
                                          num_encoder_layers=num_encoder_layers,
                                          num_decoder_layers=num_decoder_layers, dropout=dropout_p)
        self.out = nn.Linear(dim_model, latent)

    def forward(self, src, tgt, tgt_mask=None, src_pad_mask=None, tgt_pad_mask=None):
        src = self.embedding(src) * math.sqrt(self.dim_model)
        tgt = self.embedding(tgt) * math.sqrt(self.dim_model)
        src = self.positional_encoder(src)
        tgt = self.positional_encoder(tgt)
        src = src.permute(1, 0, 2)
        tgt = tgt.permute(1, 0, 2)
        y = self.transformer(src, tgt, tgt_mask=tgt_mask, src_key_padding_mask=src_pad_mask,
                             tgt_key_padding_mask=tgt_pad_mask)
        return self.out(y)

    def get_tgt_mask(self, size):
        mask = torch.tril(torch.ones(size, size) == 1).float()
        mask = mask.masked_fill(mask == 0, float("-inf"))
        mask = mask.masked_fill(mask == 1, float(0.0))
        return mask

    def create_pad_mask(self, matrix, pad_token):
        return matrix == pad_token
