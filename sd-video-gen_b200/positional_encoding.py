"""Drop-in for models/positional_encoding.py:7-35 (the table only; the add is fused into the embedding
GEMM's epilogue on the device).  Kept as a module so ``state_dict`` carries the buffer
``positional_encoder.pos_encoding`` of shape (max_len, 1, dim_model) exactly like the reference."""
import math

import torch
import torch.nn as nn


class PositionalEncoding(nn.Module):
    def __init__(self, dim_model, dropout_p, max_len):
        super().__init__()
        self.dropout = nn.Dropout(dropout_p)
        table = torch.zeros(max_len, dim_model)
        pos = torch.arange(0, max_len, dtype=torch.float).view(-1, 1)
        div = torch.exp(torch.arange(0, dim_model, 2).float() * (-math.log(10000.0)) / dim_model)
        table[:, 0::2] = torch.sin(pos * div)
        table[:, 1::2] = torch.cos(pos * div)
        self.register_buffer("pos_encoding", table.unsqueeze(0).transpose(0, 1))

    def forward(self, token_embedding):
        raise RuntimeError("PositionalEncoding is applied inside libsdvg's embedding kernel; call the Transformer")
