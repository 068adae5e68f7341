"""Task heads (reference: modules/heads/grounding_head.py:7-55, modules/heads/pretrain_head.py:8-56)."""
import torch
from torch import nn

from .. import ops
from .layers import LayerNorm, get_mlp_head
from .registry import HEADS_REGISTRY


@HEADS_REGISTRY.register()
class GroundHeadV1(nn.Module):
    def __init__(self, cfg, input_size=768, hidden_size=768, sem_cls_size=607, dropout=0.3, detach_all_aux_loss=False):
        super().__init__()
        self.og3d_head = get_mlp_head(input_size, hidden_size, 1, dropout=dropout)
        self.txt_clf_head = get_mlp_head(input_size, hidden_size, sem_cls_size, dropout=dropout)
        self.obj3d_clf_head = get_mlp_head(input_size, hidden_size, sem_cls_size, dropout=dropout)
        self.obj3d_clf_pre_head = get_mlp_head(input_size, hidden_size, sem_cls_size, dropout=dropout)
        self.detach_all_aux_loss = detach_all_aux_loss

    def forward(self, txt_embeds, obj_embeds, obj_pre_embeds, obj_masks, **kwargs):
        og3d_logits = self.og3d_head(obj_embeds).squeeze(2)
        og3d_logits = og3d_logits.masked_fill_(obj_masks.logical_not(), -float('inf'))
        if self.detach_all_aux_loss:
            txt_embeds, obj_embeds, obj_pre_embeds = txt_embeds.detach(), obj_embeds.detach(), obj_pre_embeds.detach()
        return (self.txt_clf_head(txt_embeds[:, 0]), self.obj3d_clf_head(obj_embeds),
                self.obj3d_clf_pre_head(obj_pre_embeds), og3d_logits)


@HEADS_REGISTRY.register()
class GroundHead(nn.Module):
    def __init__(self, cfg, input_size=768, hidden_size=768, dropout=0.3):
        super().__init__()
        self.og3d_head = get_mlp_head(input_size, hidden_size, 1, dropout=dropout)

    def forward(self, obj_embeds, obj_masks=None, **kwargs):
        og3d_logits = self.og3d_head(obj_embeds).squeeze(2)
        if obj_masks is not None:
            og3d_logits = og3d_logits.masked_fill_(obj_masks.logical_not(), -float('inf'))
        return og3d_logits


class _Transform(nn.Module):
    """dense -> gelu -> LayerNorm; parameter names `dense.*`, `LayerNorm.*` as in pretrain_head.py:8-19."""

    def __init__(self, width):
        super().__init__()
        self.dense, self.LayerNorm = nn.Linear(width, width), LayerNorm(width)


class BertLMPredictionHead(nn.Module):
    """BERT masked-LM head (pretrain_head.py:22-32): logits = decoder(LayerNorm(gelu(dense(h)))) + bias, with the
    state_dict keys `transform.dense.*`, `transform.LayerNorm.*`, `decoder.weight`, `bias`.  The dense+gelu and the
    decoder+bias steps go through `ops.linear` so that they pick up fused epilogues."""

    def __init__(self, hidden_size, vocab_size):
        super().__init__()
        self.transform = _Transform(hidden_size)
        self.decoder = nn.Linear(hidden_size, vocab_size, bias=False)
        self.bias = nn.Parameter(torch.zeros(vocab_size))

    def forward(self, hidden_states):
        t = self.transform
        h = t.LayerNorm(ops.linear(hidden_states, t.dense.weight, t.dense.bias, activation="gelu"))
        return ops.padded_vocab_linear(h, self.decoder.weight, self.bias)


def _lm_heads(owner, hidden_size, **vocabs):
    for name, size in vocabs.items():
        setattr(owner, name, BertLMPredictionHead(hidden_size, size))


@HEADS_REGISTRY.register()
class PretrainHeadV1(nn.Module):
    def __init__(self, cfg, hidden_size=768, vocab_size=30522):
        super().__init__()
        _lm_heads(self, hidden_size, lm_pred_head=vocab_size)

    def forward(self, txt_embeds, **kwargs):
        return self.lm_pred_head(txt_embeds)


@HEADS_REGISTRY.register()
class OVPretrainHead(nn.Module):
    def __init__(self, cfg, hidden_size=768, vocab_size=30522, obj_vocab_size=607):
        super().__init__()
        _lm_heads(self, hidden_size, lm_pred_head=vocab_size, obj_pred_head=obj_vocab_size)

    def forward(self, txt_embeds, obj_embeds, **kwargs):
        return self.lm_pred_head(txt_embeds), self.obj_pred_head(obj_embeds)
