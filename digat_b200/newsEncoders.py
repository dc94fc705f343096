"""Drop-in for the MSA news (title) encoder of reference newsEncoders.py:58-82 (layers.py:50-115) on the sm_100a kernels
(SURVEY.md section 8(f) row 4): same class / parameter names and shapes (``word_embedding.weight``,
``multiheadSelfattention.W_{Q,K,V}``, ``attention.affine{1,2}``), same ``forward(title_text, title_mask)``.

In the reference's evaluation this encoder runs once over all news (util.compute_scores, util.py:24-33) and its output is the
cached embedding table the graph encoder reads: that is the no-grad path below.  With gradients enabled (reference
model.py:68-69 trains it end to end) ``forward`` runs the same kernels as autograd nodes with their backward kernels
(digat_msa_attention_bwd, digat_additive_pool_bwd, digat_scatter_add_rows; projections through autograd_ops.lin).

  word embeddings    digat_gather_rows_i32            [titles*T, E]
  Q | K | V          ONE projection GEMM              [titles*T, 3*h*dk]   (stacked weight, biases of Q and V)
  attention + relu   digat_msa_attention_fwd          one warp per (title, head), lane = query token
  affine1            projection GEMM                  [titles*T, A]
  tanh . w2, masked softmax over tokens, weighted sum   digat_additive_pool_fwd
The CNN encoder (newsEncoders.py:27-55; cnn_method 'naive' and 'group3') is an im2col gather + ONE GEMM with a relu epilogue
+ the same pooling (class CNN below)."""
import math
import os
import pickle

import torch
import torch.nn as nn

from . import _lib
from .graphEncoders import PackedWeight, _ptr, _stream, linear


# ------------------------------------------------------------------------------------------------ training path
# Autograd nodes over the backward kernels of news_encoder.cuh; the projections go through autograd_ops.lin (tcgen05 /
# exact-fp32 GEMMs with their dgrad and wgrad), dropout masks come from torch like in the graph encoder's training path.
class _EmbeddingFn(torch.autograd.Function):
    """rows = table[idx] (digat_gather_rows_i32, optionally into a column block of a wider matrix); the backward scatter-adds
    into a zeroed table-shaped gradient (digat_scatter_add_rows, float atomics like torch's embedding backward)."""

    @staticmethod
    def forward(ctx, table, idx, err):
        rows, E = idx.numel(), table.shape[1]
        out = torch.empty((rows, E), device=table.device, dtype=torch.float32)
        _lib.call('digat_gather_rows_i32', table.data_ptr(), table.shape[0], idx.data_ptr(), out.data_ptr(), E, rows, E,
                  err.data_ptr(), _stream())
        ctx.save_for_backward(idx)
        ctx.shape = tuple(table.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        dout = dout.contiguous()
        dtable = torch.zeros(ctx.shape, device=dout.device, dtype=torch.float32)
        _lib.call('digat_scatter_add_rows', dtable.data_ptr(), ctx.shape[0], idx.data_ptr(), dout.data_ptr(), dout.stride(0),
                  idx.numel(), ctx.shape[1], _stream())
        return dtable, None, None


class _MsaAttentionFn(torch.autograd.Function):
    """H = relu(multi-head self-attention) from the stacked Q | K | V rows (digat_msa_attention_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, qkv, n_titles, T, heads, dk):
        qkv = qkv.contiguous()
        hd = heads * dk
        H = torch.empty((n_titles * T, hd), device=qkv.device, dtype=torch.float32)
        _lib.call('digat_msa_attention_fwd', qkv.data_ptr(), qkv.stride(0), H.data_ptr(), hd, n_titles, T, heads, dk, _stream())
        ctx.save_for_backward(qkv, H)
        ctx.dims = (n_titles, T, heads, dk)
        return H

    @staticmethod
    def backward(ctx, dH):
        qkv, H = ctx.saved_tensors
        n_titles, T, heads, dk = ctx.dims
        dH = dH.contiguous()
        dqkv = torch.empty_like(qkv)
        _lib.call('digat_msa_attention_bwd', qkv.data_ptr(), qkv.stride(0), H.data_ptr(), H.stride(0), dH.data_ptr(), dH.stride(0),
                  dqkv.data_ptr(), dqkv.stride(0), n_titles, T, heads, dk, _stream())
        return dqkv, None, None, None, None


class _AdditivePoolFn(torch.autograd.Function):
    """out[title] = sum_t softmax_t(mask(tanh(att_pre_t) . w2)) H_t  (digat_additive_pool_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, att_pre, w2, H, mask, n_titles, T):
        att_pre, w2, H = att_pre.contiguous(), w2.contiguous(), H.contiguous()
        D, A = H.shape[1], att_pre.shape[1]
        out = torch.empty((n_titles, D), device=H.device, dtype=torch.float32)
        _lib.call('digat_additive_pool_fwd', att_pre.data_ptr(), A, w2.data_ptr(), H.data_ptr(), D, mask.data_ptr(),
                  out.data_ptr(), D, n_titles, T, A, D, _stream())
        ctx.save_for_backward(att_pre, w2, H, mask)
        ctx.dims = (n_titles, T)
        return out

    @staticmethod
    def backward(ctx, dout):
        from .autograd_ops import colsum
        att_pre, w2, H, mask = ctx.saved_tensors
        n_titles, T = ctx.dims
        D, A = H.shape[1], att_pre.shape[1]
        dout = dout.contiguous()
        dH = torch.empty_like(H)
        datt = torch.empty_like(att_pre)
        dw2_part = torch.empty((n_titles, A), device=H.device, dtype=torch.float32)
        _lib.call('digat_additive_pool_bwd', att_pre.data_ptr(), A, w2.data_ptr(), H.data_ptr(), D, mask.data_ptr(),
                  dout.data_ptr(), dout.stride(0), dH.data_ptr(), D, datt.data_ptr(), A, dw2_part.data_ptr(), n_titles, T, A, D,
                  _stream())
        return datt, colsum(dw2_part), dH, None, None, None


def _train_prologue(enc, title_text, title_mask):
    if not title_text.is_cuda:
        raise RuntimeError('title_text must be a CUDA tensor (digat_b200 has no CPU fallback)')
    dev = title_text.device
    _lib.require_device(dev.index if dev.index is not None else torch.cuda.current_device())
    B, news_num, T = title_text.shape
    if T != enc.max_sentence_length:
        raise RuntimeError('title length %d != max_title_length %d' % (T, enc.max_sentence_length))
    n_titles = B * news_num
    mask = title_mask.reshape(n_titles, T)
    mask = (mask != 0).contiguous() if mask.dtype != torch.bool else mask.contiguous()
    return B, news_num, T, n_titles, mask, torch.zeros(1, dtype=torch.int32, device=dev)


def _wants_grad(enc):
    return torch.is_grad_enabled() and any(p.requires_grad for p in enc.parameters())


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
        if _wants_grad(self):
            return self._forward_train(title_text, title_mask)
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


def _msa_forward_train(self, title_text, title_mask):
    """newsEncoders.py:71-82 with autograd: dropout on the word embeddings (train mode), Q | K | V as ONE projection, the
    attention and pooling kernels with their backward kernels; gradients reach every parameter incl. the embedding table."""
    from .autograd_ops import lin
    import torch.nn.functional as F
    B, news_num, T, n_titles, mask, err = _train_prologue(self, title_text, title_mask)
    m = self.multiheadSelfattention
    tok = title_text.reshape(-1).to(torch.int32).contiguous()
    emb = _EmbeddingFn.apply(self.word_embedding.weight, tok, err)
    if self.training and self.dropout_rate > 0:
        emb = F.dropout(emb, self.dropout_rate, True)
    qkv_W = torch.cat([m.W_Q.weight, m.W_K.weight, m.W_V.weight], 0)
    qkv_b = torch.cat([m.W_Q.bias, torch.zeros_like(m.W_Q.bias), m.W_V.bias], 0)
    pad = (-emb.shape[1]) % 16
    if pad and emb.shape[0] > 2048:
        # E = 300 is not a multiple of 16: the dgrad (N = E) and the split-K wgrad (K = E) of this projection would fall off the
        # tensor-core path onto the exact-fp32 CUDA-core kernels (2 x 147 GFLOP per step at MIND sizes: 18 of 36 ms).  Zero
        # columns keep the product unchanged; autograd slices their gradients off again.
        emb, qkv_W = F.pad(emb, (0, pad)), F.pad(qkv_W, (0, pad))
    qkv = lin(emb, qkv_W, qkv_b)
    H = _MsaAttentionFn.apply(qkv, n_titles, T, m.h, m.d_k)
    att = lin(H, self.attention.affine1.weight, self.attention.affine1.bias)
    out = _AdditivePoolFn.apply(att, self.attention.affine2.weight.reshape(-1), H, mask, n_titles, T)
    if int(err.item()) != 0:
        raise RuntimeError('token id out of range in title_text (reference: nn.Embedding raises)')
    return out.view(B, news_num, self.news_embedding_dim)


MSA._forward_train = _msa_forward_train


class Conv1D(nn.Module):
    """Parameter container with the names of reference layers.py:7-26 ('naive' and 'group3'; 'group5' cannot run in the
    reference either: layers.py:43-48 concatenates its padding column along the channel dimension)."""

    def __init__(self, cnn_method, in_channels, cnn_kernel_num, cnn_window_size):
        super().__init__()
        if cnn_method not in ('naive', 'group3'):
            raise Exception("cnn_method %r is not implemented on the sm_100a path ('naive' and 'group3' are)" % cnn_method)
        self.cnn_method, self.in_channels = cnn_method, in_channels
        if cnn_method == 'naive':
            if cnn_window_size % 2 != 1:
                raise Exception('cnn_window_size must be odd (same-length output, reference padding (window-1)//2)')
            self.conv = nn.Conv1d(in_channels, cnn_kernel_num, kernel_size=cnn_window_size, padding=(cnn_window_size - 1) // 2)
            self.window = cnn_window_size
        else:
            assert cnn_kernel_num % 3 == 0
            self.conv1 = nn.Conv1d(in_channels, cnn_kernel_num // 3, kernel_size=1, padding=0)
            self.conv2 = nn.Conv1d(in_channels, cnn_kernel_num // 3, kernel_size=3, padding=1)
            self.conv3 = nn.Conv1d(in_channels, cnn_kernel_num // 3, kernel_size=5, padding=2)
            self.window = 5

    def initialize(self):
        pass

    def packed(self):
        """One [kernels, window * E] matrix acting on the im2col rows [x[t-pad] | ... | x[t+pad]] (+ its bias): a window-w
        convolution occupies the w central column blocks, zero elsewhere -- the 'group3' concatenation is ONE product."""
        convs = [self.conv] if self.cnn_method == 'naive' else [self.conv1, self.conv2, self.conv3]
        E, W = self.in_channels, self.window
        rows = []
        for c in convs:
            w = c.weight.detach().float()                      # [out, E, k]
            k = w.shape[2]
            full = w.new_zeros((w.shape[0], W, E))
            lo = (W - k) // 2
            full[:, lo:lo + k, :] = w.permute(0, 2, 1)
            rows.append(full.reshape(w.shape[0], W * E))
        return torch.cat(rows, 0).contiguous(), torch.cat([c.bias.detach().float() for c in convs], 0).contiguous()


class CNN(NewsEncoder):
    """Drop-in for the CNN news encoder of reference newsEncoders.py:27-55 (no-grad path; training: _cnn_forward_train):
      word embeddings   digat_gather_rows_i32, once per window offset, straight into the column blocks of the im2col matrix
                        [titles*T, window*E] (positions outside the title read an appended zero row = Conv1d's zero padding)
      conv + relu       ONE projection GEMM against the packed kernels (relu in its epilogue)
      attention pooling affine1 GEMM + digat_additive_pool_fwd, as in MSA."""

    def __init__(self, config):
        super().__init__(config)
        self.max_sentence_length = config.max_title_length
        self.cnn_kernel_num = config.cnn_kernel_num
        self.conv = Conv1D(config.cnn_method, config.word_embedding_dim, config.cnn_kernel_num, config.cnn_window_size)
        self.news_embedding_dim = config.cnn_kernel_num
        self.attention = Attention(self.news_embedding_dim, config.attention_dim)
        if self.news_embedding_dim % 4 != 0 or config.word_embedding_dim % 4 != 0:
            raise Exception('word_embedding_dim and cnn_kernel_num must be multiples of 4 for the sm_100a kernels')
        self._packed = self._packed_key = None

    def initialize(self):
        super().initialize()
        self.conv.initialize()
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
            raise RuntimeError('CNN parameters must live on a CUDA device (digat_b200 has no CPU fallback)')
        _lib.require_device(dev.index if dev.index is not None else torch.cuda.current_device())
        with torch.no_grad():
            table = self.word_embedding.weight.detach().float()
            conv_W, conv_b = self.conv.packed()
            w = {'table': torch.cat([table, table.new_zeros((1, table.shape[1]))], 0).contiguous(),   # + the zero-padding row
                 'conv_W': PackedWeight(conv_W), 'conv_b': conv_b,
                 'a1_W': PackedWeight(self.attention.affine1.weight.detach().float().contiguous()),
                 'a1_b': self.attention.affine1.bias.detach().float().contiguous(),
                 'w2': self.attention.affine2.weight.detach().float().reshape(-1).contiguous()}
        self._packed, self._packed_key = w, key
        return w

    def forward(self, title_text, title_mask):
        """title_text [B, news_num, T] integer token ids, title_mask [B, news_num, T] -> [B, news_num, cnn_kernel_num]."""
        if _wants_grad(self):
            return self._forward_train(title_text, title_mask)
        if not title_text.is_cuda:
            raise RuntimeError('title_text must be a CUDA tensor (digat_b200 has no CPU fallback)')
        w = self._weights()
        B, news_num, T = title_text.shape
        if T != self.max_sentence_length:
            raise RuntimeError('title length %d != max_title_length %d' % (T, self.max_sentence_length))
        n_titles, E, Fk, W = B * news_num, self.word_embedding_dim, self.news_embedding_dim, self.conv.window
        pad, V = (W - 1) // 2, w['table'].shape[0] - 1
        dev = title_text.device
        with torch.no_grad():
            tok = title_text.reshape(n_titles, T).to(torch.int32)
            mask = title_mask.reshape(n_titles, T)
            mask = (mask != 0).contiguous() if mask.dtype != torch.bool else mask.contiguous()
            err = torch.zeros(1, dtype=torch.int32, device=dev)
            bad = ((tok < 0) | (tok >= V)).any()                          # the appended zero row must not be reachable by a token
            padded = torch.nn.functional.pad(tok, (pad, pad), value=V)    # index V = the zero row
            X = torch.empty((n_titles * T, W * E), device=dev, dtype=torch.float32)
            for k in range(W):
                idx = padded[:, k:k + T].contiguous()
                _lib.call('digat_gather_rows_i32', w['table'].data_ptr(), V + 1, idx.data_ptr(), X.data_ptr() + 4 * k * E, W * E,
                          n_titles * T, E, err.data_ptr(), _stream())
            H = linear(X, w['conv_W'], w['conv_b'], relu=True)                           # [titles*T, kernels]
            att = linear(H, w['a1_W'], w['a1_b'])                                        # tanh in the pooling kernel
            out = torch.empty((n_titles, Fk), device=dev, dtype=torch.float32)
            _lib.call('digat_additive_pool_fwd', att.data_ptr(), att.stride(0), w['w2'].data_ptr(), H.data_ptr(), Fk,
                      mask.data_ptr(), out.data_ptr(), Fk, n_titles, T, att.shape[1], Fk, _stream())
            if int(err.item()) != 0 or bool(bad):
                raise RuntimeError('token id out of range in title_text (reference: nn.Embedding raises)')
        return out.view(B, news_num, Fk)


def _conv_packed_train(conv):
    """Conv1D.packed() built from the live parameters with differentiable torch ops (permute / pad / cat)."""
    convs = [conv.conv] if conv.cnn_method == 'naive' else [conv.conv1, conv.conv2, conv.conv3]
    E, W = conv.in_channels, conv.window
    rows = []
    for c in convs:
        k = c.weight.shape[2]
        lo = (W - k) // 2
        full = torch.nn.functional.pad(c.weight.permute(0, 2, 1), (0, 0, lo, W - k - lo))      # [out, W, E], zero blocks outside
        rows.append(full.reshape(c.weight.shape[0], W * E))
    return torch.cat(rows, 0), torch.cat([c.bias for c in convs], 0)


def _cnn_forward_train(self, title_text, title_mask):
    """newsEncoders.py:41-54 with autograd: dropout on the word embeddings and on the relu'd convolution output (train mode)."""
    from .autograd_ops import lin
    import torch.nn.functional as F
    B, news_num, T, n_titles, mask, err = _train_prologue(self, title_text, title_mask)
    W, V = self.conv.window, self.word_embedding.weight.shape[0]
    pad = (W - 1) // 2
    tok = title_text.reshape(n_titles, T).to(torch.int32)
    if bool(((tok < 0) | (tok >= V)).any()):
        raise RuntimeError('token id out of range in title_text (reference: nn.Embedding raises)')
    emb = _EmbeddingFn.apply(self.word_embedding.weight, tok.reshape(-1).contiguous(), err)
    if self.training and self.dropout_rate > 0:
        emb = F.dropout(emb, self.dropout_rate, True)
    # im2col [titles*T, W*E]: token t's row holds x[t-pad] | ... | x[t+pad], zeros outside the title (Conv1d's zero padding)
    x = F.pad(emb.view(n_titles, T, -1), (0, 0, pad, pad))
    X = torch.cat([x[:, k:k + T, :] for k in range(W)], 2).reshape(n_titles * T, -1)
    conv_W, conv_b = _conv_packed_train(self.conv)
    H = torch.relu(lin(X, conv_W, conv_b))
    if self.training and self.dropout_rate > 0:
        H = F.dropout(H, self.dropout_rate, True)
    att = lin(H, self.attention.affine1.weight, self.attention.affine1.bias)
    out = _AdditivePoolFn.apply(att, self.attention.affine2.weight.reshape(-1), H, mask, n_titles, T)
    return out.view(B, news_num, self.news_embedding_dim)


CNN._forward_train = _cnn_forward_train
