"""Drop-in for the MSA news (title) encoder of reference newsEncoders.py:58-82 (layers.py:50-115) on the sm_100a kernels
(SURVEY.md section 8(f) row 4): same class / parameter names and shapes (``word_embedding.weight``,
``multiheadSelfattention.W_{Q,K,V}``, ``attention.affine{1,2}``), same ``forward(title_text, title_mask)``.

Inference only (``torch.no_grad``): in the reference's evaluation this encoder runs once over all news
(util.compute_scores, util.py:24-33) and its output is the cached embedding table the graph encoder reads.  Training it
(reference model.py:68-69) needs its backward kernels, which are not written: asking for gradients raises.

  word embeddings    digat_gather_rows_i32            [titles*T, E]
  Q | K | V          ONE projection GEMM              [titles*T, 3*h*dk]   (stacked weight, biases of Q and V)
  attention + relu   digat_msa_attention_fwd          one warp per (title, head), lane = query token
  affine1            projection GEMM                  [titles*T, A]
  tanh . w2, masked softmax over tokens, weighted sum   digat_additive_pool_fwd
The CNN encoder (newsEncoders.py:27-55) is not implemented (the reference's default and published setting is MSA)."""
import math
import os
import pickle

import torch
import torch.nn as nn

from . import _lib
from .graphEncoders import PackedWeight, _ptr, _stream, linear


class MultiHeadAttention(nn.Module):
    """Parameter container with the names of reference layers.py:50-65."""

    def __init__(self, h, d_model, len_q, len_k, d_k, d_v):
        super().__init__()
        self.h, self.d_model, self.len_q, self.len_k, self.d_k, self.d_v = h, d_model, len_q, len_k, d_k, d_v
        self.out_dim = h * d_v
        self.attention_scalar = math.sqrt(float(d_k))
        self.W_K = nn.Linear(d_model, h * d_k, bias=False)
        self.W_Q = nn.Linear(d_model, h * d_k, bias=True)
        self.W_V = nn.Linear(d_model, h * d_v, bias=True)

    def initialize(self):
        nn.init.zeros_(self.W_Q.bias)
        nn.init.zeros_(self.W_V.bias)


class Attention(nn.Module):
    """Parameter container with the names of reference layers.py:98-106."""

    def __init__(self, feature_dim, attention_dim):
        super().__init__()
        self.affine1 = nn.Linear(feature_dim, attention_dim, bias=True)
        self.affine2 = nn.Linear(attention_dim, 1, bias=False)

    def initialize(self):
        nn.init.xavier_uniform_(self.affine1.weight, gain=nn.init.calculate_gain('tanh'))
        nn.init.zeros_(self.affine1.bias)
        nn.init.xavier_uniform_(self.affine2.weight)


class NewsEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.word_embedding_dim = config.word_embedding_dim
        self.word_embedding = nn.Embedding(num_embeddings=config.vocabulary_size, embedding_dim=self.word_embedding_dim)
        # the reference reads its preprocessed embedding matrix here (newsEncoders.py:14-15); same file name when present
        path = 'word_embedding-%s-%s-%s-%s.pkl' % (getattr(config, 'word_threshold', 3), config.word_embedding_dim,
                                                   config.max_title_length, getattr(config, 'dataset', 'MIND-small'))
        if os.path.isfile(path):
            with open(path, 'rb') as f:
                self.word_embedding.weight.data.copy_(pickle.load(f))
        self.dropout_rate = float(config.dropout_rate)

    def initialize(self):
        pass

    def forward(self, title_text, title_mask):
        raise Exception('Function forward must be implemented at sub-class')


class MSA(NewsEncoder):
    def __init__(self, config):
        super().__init__(config)
        self.max_sentence_length = config.max_title_length
        self.multiheadSelfattention = MultiHeadAttention(config.MSA_head_num, config.word_embedding_dim,
                                                         config.max_title_length, config.max_title_length,
                                                         config.MSA_head_dim, config.MSA_head_dim)
        self.news_embedding_dim = config.MSA_head_num * config.MSA_head_dim
        self.attention = Attention(self.news_embedding_dim, config.attention_dim)
        if self.news_embedding_dim % 4 != 0 or config.word_embedding_dim % 4 != 0:
            raise Exception('word_embedding_dim and MSA_head_num * MSA_head_dim must be multiples of 4 for the sm_100a kernels')
        self._packed = self._packed_key = None

    def initialize(self):
        super().initialize()
        self.multiheadSelfattention.initialize()
        self.attention.initialize()

    def invalidate_packed(self):
        self._packed = self._packed_key = None

    def train(self, mode: bool = True):
        self.invalidate_packed()
        return super().train(mode)

    def _weights(self):
        params = list(self.parameters())
        key = tuple((p.data_ptr(), p._version) for p in params)
        if self._packed is not None and key == self._packed_key:
            return self._packed
        dev = params[0].device
        if dev.type != 'cuda':
            raise RuntimeError('MSA parameters must live on a CUDA device (digat_b200 has no CPU fallback)')
        _lib.require_device(dev.index if dev.index is not None else torch.cuda.current_device())
        m = self.multiheadSelfattention
        with torch.no_grad():
            w = {'table': self.word_embedding.weight.detach().float().contiguous(),
                 'qkv_W': PackedWeight(torch.cat([m.W_Q.weight, m.W_K.weight, m.W_V.weight], 0).float().contiguous()),
                 'qkv_b': torch.cat([m.W_Q.bias, torch.zeros_like(m.W_Q.bias), m.W_V.bias], 0).float().contiguous(),
                 'a1_W': PackedWeight(self.attention.affine1.weight.detach().float().contiguous()),
                 'a1_b': self.attention.affine1.bias.detach().float().contiguous(),
                 'w2': self.attention.affine2.weight.detach().float().reshape(-1).contiguous()}
        self._packed, self._packed_key = w, key
        return w

    def forward(self, title_text, title_mask):
        """title_text [B, news_num, T] integer token ids, title_mask [B, news_num, T] -> [B, news_num, h*dk]."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError('MSA news encoder: the sm_100a path is inference-only (run under torch.no_grad())')
        if not title_text.is_cuda:
            raise RuntimeError('title_text must be a CUDA tensor (digat_b200 has no CPU fallback)')
        w = self._weights()
        B, news_num, T = title_text.shape
        if T != self.max_sentence_length:
            raise RuntimeError('title length %d != max_title_length %d' % (T, self.max_sentence_length))
        m = self.multiheadSelfattention
        n_titles, E, hd = B * news_num, self.word_embedding_dim, self.news_embedding_dim
        dev = title_text.device
        with torch.no_grad():
            tok = title_text.reshape(-1).to(torch.int32).contiguous()
            mask = title_mask.reshape(n_titles, T)
            mask = (mask != 0).contiguous() if mask.dtype != torch.bool else mask.contiguous()
            err = torch.zeros(1, dtype=torch.int32, device=dev)
            emb = torch.empty((n_titles * T, E), device=dev, dtype=torch.float32)
            _lib.call('digat_gather_rows_i32', w['table'].data_ptr(), w['table'].shape[0], tok.data_ptr(), emb.data_ptr(), E,
                      n_titles * T, E, err.data_ptr(), _stream())
            qkv = linear(emb, w['qkv_W'], w['qkv_b'])                                    # [titles*T, 3*hd]
            H = torch.empty((n_titles * T, hd), device=dev, dtype=torch.float32)
            _lib.call('digat_msa_attention_fwd', qkv.data_ptr(), qkv.stride(0), H.data_ptr(), hd, n_titles, T, m.h, m.d_k,
                      _stream())
            att = linear(H, w['a1_W'], w['a1_b'])                                        # [titles*T, A] (tanh in the pooling kernel)
            out = torch.empty((n_titles, hd), device=dev, dtype=torch.float32)
            _lib.call('digat_additive_pool_fwd', att.data_ptr(), att.stride(0), w['w2'].data_ptr(), H.data_ptr(), hd,
                      mask.data_ptr(), out.data_ptr(), hd, n_titles, T, att.shape[1], hd, _stream())
            if int(err.item()) != 0:
                raise RuntimeError('token id out of range in title_text (reference: nn.Embedding raises)')
        return out.view(B, news_num, hd)


class CNN(NewsEncoder):
    def __init__(self, config):
        raise Exception('CNN news encoder is not implemented on the sm_100a path (use --news_encoder=MSA)')
