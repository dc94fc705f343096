"""Drop-in for reference model.py: the ``graph_encoder`` string dispatch (model.py:18-31), ``Model.inference``
(model.py:87-90), ``Model.forward`` (model.py:54-77) and its logits on top of the sm_100a encoders.

The news (text) encoder is outside the north-star hot path: the scoring path works on news
embeddings, exactly as reference util.compute_scores does after caching them (util.py:24-33) -- ``forward_embeddings`` is
``forward`` from the point where the embeddings exist (and is what trains, through autograd_ops).  When the config carries
the text-side fields (``vocabulary_size`` ...) the MSA news encoder (newsEncoders.py) is built too and
``forward`` takes the reference's token tensors (with gradients enabled both encoders run their training paths)."""
import torch
import torch.nn as nn

from . import _lib, graphEncoders


class Model(nn.Module):
    def __init__(self, config, news_embedding_dim: int = 400):
        super().__init__()
        from .ablation_encoders import ENCODERS                                 # the string dispatch of model.py:18-31
        if config.graph_encoder not in ENCODERS:
            raise Exception(config.graph_encoder + ' is not implemented')      # same wording as model.py:31
        self.news_encoder = None
        if hasattr(config, 'vocabulary_size'):                                  # text side available: model.py:11-16
            from . import newsEncoders
            kind = getattr(config, 'news_encoder', 'MSA')           # reference model.py:10-15
            if kind == 'MSA':
                self.news_encoder = newsEncoders.MSA(config)
            elif kind == 'CNN':
                self.news_encoder = newsEncoders.CNN(config)
            else:
                raise Exception(kind + ' is not implemented')
            news_embedding_dim = self.news_encoder.news_embedding_dim
            self.max_title_length = config.max_title_length
        self.graph_encoder = ENCODERS[config.graph_encoder](config, news_embedding_dim)
        self.model_name = getattr(config, 'news_encoder', 'MSA') + '-' + config.graph_encoder
        self.max_history_num = config.max_history_num
        self.category_num = config.category_num + 1
        self.news_embedding_dim = news_embedding_dim
        self.representation_dim = news_embedding_dim
        self.news_graph_size = config.news_graph_size
        self.user_graph_size = config.max_history_num + config.category_num

    def initialize(self):
        if self.news_encoder is not None:
            self.news_encoder.initialize()
        self.graph_encoder.initialize()

    def forward(self, user_title_text, user_title_mask, user_graph, user_category_mask, user_category_indices,
                news_title_text, news_title_mask, news_graph, news_graph_mask):
        """Reference Model.forward (model.py:54-77) on token tensors; needs the news encoder.  With gradients enabled both encoders
        run their training paths (end-to-end, as reference trainer.py)."""
        if self.news_encoder is None:
            raise RuntimeError('Model was built without the text-side config fields: use forward_embeddings')
        bs, news_num = news_graph.shape[0], news_graph.shape[1]
        T = self.max_title_length
        cand = self.news_encoder(news_title_text.reshape(bs * news_num, self.news_graph_size, T),
                                 news_title_mask.reshape(bs * news_num, self.news_graph_size, T))
        hist = self.news_encoder(user_title_text, user_title_mask)
        return self.forward_embeddings(hist, user_graph, user_category_mask, user_category_indices,
                                       cand.view(bs, news_num, self.news_graph_size, self.news_embedding_dim),
                                       news_graph, news_graph_mask)

    def inference(self, user_news_embedding, user_graph, user_category_mask, user_category_indices,
                  candidate_news_embedding, news_graph, news_graph_mask, c_n0):
        """Same argument order as reference Model.inference (model.py:87); returns logits [batch_size]."""
        news_repr, user_repr = self.graph_encoder.inference(candidate_news_embedding, news_graph, news_graph_mask,
                                                            user_news_embedding, user_graph, user_category_mask,
                                                            user_category_indices, c_n0)
        return logits(news_repr, user_repr)

    def forward_embeddings(self, user_news_embedding, user_graph, user_category_mask, user_category_indices,
                           candidate_news_embedding, news_graph, news_graph_mask):
        """Model.forward (model.py:54-77) from the point where the news encoder has produced embeddings:
        user tensors [bs,...] are expanded over the news_num candidates of candidate_news_embedding
        [bs, news_num, n_n, D]; returns logits [bs, news_num]."""
        bs, news_num = news_graph.shape[0], news_graph.shape[1]
        bn = bs * news_num
        ex = lambda t: t.unsqueeze(1).expand(-1, news_num, *([-1] * (t.dim() - 1))).reshape(bn, *t.shape[1:])
        news_repr, user_repr = self.graph_encoder(
            candidate_news_embedding.reshape(bn, self.news_graph_size, self.news_embedding_dim),
            news_graph.reshape(bn, self.news_graph_size, self.news_graph_size),
            news_graph_mask.reshape(bn, self.news_graph_size),
            ex(user_news_embedding), ex(user_graph), ex(user_category_mask), ex(user_category_indices))
        if news_repr.requires_grad or user_repr.requires_grad:
            return (user_repr * news_repr).sum(dim=1).view(bs, news_num)
        return logits(news_repr, user_repr).view(bs, news_num)


def logits(news_repr, user_repr):
    B, D = news_repr.shape
    out = torch.empty(B, device=news_repr.device, dtype=torch.float32)
    _lib.call('digat_logits', news_repr.data_ptr(), user_repr.data_ptr(), out.data_ptr(), B, D,
              torch.cuda.current_stream().cuda_stream)
    return out
