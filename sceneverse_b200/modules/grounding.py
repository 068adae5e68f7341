"""Language-object fusion encoders (reference: modules/grounding/unified_encoder.py:12-177)."""
import torch
from torch import nn

from .layers import (LayerNorm, TransformerDecoderLayer, TransformerEncoderLayer, TransformerSpatialDecoderLayer,
                     calc_pairwise_locs, init_weights_bert, layer_repeat)
from .registry import GROUNDING_REGISTRY


def _loc_layers(dim_loc, hidden_size):
    return layer_repeat(nn.Sequential(nn.Linear(dim_loc, hidden_size), LayerNorm(hidden_size)), 1)


@GROUNDING_REGISTRY.register()
class UnifiedSpatialCrossEncoderV2(nn.Module):
    """unified_encoder.py:121-177: 4 x joint self-attention over [text ; objects] with location and
    token-type embeddings re-added before every layer (the reference's .cuda() calls are replaced by
    device-of-input placement)."""

    def __init__(self, cfg, hidden_size=768, dim_feedforward=2048, num_attention_heads=12, num_layers=4, dim_loc=6):
        super().__init__()
        self.unified_encoder = layer_repeat(
            TransformerEncoderLayer(hidden_size, num_attention_heads, dim_feedforward=dim_feedforward), num_layers)
        self.loc_layers = _loc_layers(dim_loc, hidden_size)
        self.token_type_embeddings = nn.Embedding(2, hidden_size)
        self.apply(init_weights_bert)

    def forward(self, txt_embeds, txt_masks, obj_embeds, obj_locs, obj_masks, output_attentions=False,
                output_hidden_states=False, **kwargs):
        txt_len, obj_len = txt_embeds.shape[1], obj_embeds.shape[1]
        tt = self.token_type_embeddings.weight
        key_padding = torch.cat((txt_masks, obj_masks), dim=1).logical_not()
        # The reference splits the joint sequence after every layer, re-adds loc_layers[0](obj_locs) + tt[1] to the objects
        # and tt[0] to the text and concatenates again (unified_encoder.py:163-172).  The re-added term is the same tensor
        # every layer (same module, same input, no dropout inside), so it is built once as one [B, T + O, C] addend and the
        # joint sequence stays joint: one add per layer instead of two adds, a cat and a split (and their backward copies).
        dt = txt_embeds.dtype
        addend = torch.cat((tt[0].to(dt).expand(txt_embeds.shape[0], txt_len, -1),
                            (self.loc_layers[0](obj_locs) + tt[1]).to(dt)), dim=1)
        joint = torch.cat((txt_embeds, obj_embeds.to(dt)), dim=1)
        for layer in self.unified_encoder:
            joint, _ = layer(joint + addend, tgt_key_padding_mask=key_padding)
        txt_embeds, obj_embeds = torch.split(joint, [txt_len, obj_len], dim=1)
        return txt_embeds, obj_embeds


@GROUNDING_REGISTRY.register()
class UnifiedSpatialCrossEncoderV1(nn.Module):
    """unified_encoder.py:60-118: objects attend to text (spatial decoder) and text attends to objects, 4 layers.
    calc_pairwise_locs is called with the reference's defaults here (spatial_dist_norm=True, spatial_dim=5)."""

    def __init__(self, cfg, hidden_size=768, num_attention_heads=12, spatial_dim=5, num_layers=4, dim_loc=6,
                 pairwise_rel_type='center'):
        super().__init__()
        pc_layer = TransformerSpatialDecoderLayer(hidden_size, num_attention_heads, dim_feedforward=2048, dropout=0.1,
                                                  activation='gelu', spatial_dim=spatial_dim, spatial_multihead=True,
                                                  spatial_attn_fusion='cond')
        self.pc_encoder = layer_repeat(pc_layer, num_layers)
        self.lang_encoder = layer_repeat(TransformerDecoderLayer(hidden_size, num_attention_heads), num_layers)
        self.loc_layers = _loc_layers(dim_loc, hidden_size)
        self.pairwise_rel_type, self.spatial_dim, self.spatial_dist_norm = pairwise_rel_type, spatial_dim, True
        self.apply(init_weights_bert)

    def forward(self, txt_embeds, txt_masks, obj_embeds, obj_locs, obj_masks, output_attentions=False,
                output_hidden_states=False, **kwargs):
        pairwise_locs = calc_pairwise_locs(obj_locs[:, :, :3], obj_locs[:, :, 3:], pairwise_rel_type=self.pairwise_rel_type)
        obj_pad, txt_pad = obj_masks.logical_not(), txt_masks.logical_not()
        loc = self.loc_layers[0](obj_locs).to(obj_embeds.dtype)   # identical every layer: computed once
        for pc_layer, lang_layer in zip(self.pc_encoder, self.lang_encoder):
            obj_embeds = obj_embeds + loc
            obj_out, _, _ = pc_layer(obj_embeds, txt_embeds, pairwise_locs, tgt_key_padding_mask=obj_pad,
                                     memory_key_padding_mask=txt_pad)
            txt_out, _, _ = lang_layer(txt_embeds, obj_embeds, tgt_key_padding_mask=txt_pad,
                                       memory_key_padding_mask=obj_pad)
            obj_embeds, txt_embeds = obj_out, txt_out
        return txt_embeds, obj_embeds


@GROUNDING_REGISTRY.register()
class EntitySpatialCrossEncoder(nn.Module):
    """unified_encoder.py:12-57: objects attend to (fixed) text only."""

    def __init__(self, cfg, hidden_size=768, num_attention_heads=12, spatial_dim=5, num_layers=4, dim_loc=6,
                 pairwise_rel_type='center'):
        super().__init__()
        layer = TransformerSpatialDecoderLayer(hidden_size, num_attention_heads, dim_feedforward=2048, dropout=0.1,
                                               activation='gelu', spatial_dim=spatial_dim, spatial_multihead=True,
                                               spatial_attn_fusion='cond')
        self.layers = layer_repeat(layer, num_layers)
        self.loc_layers = _loc_layers(dim_loc, hidden_size)
        self.pairwise_rel_type, self.spatial_dim, self.spatial_dist_norm = pairwise_rel_type, spatial_dim, True
        self.apply(init_weights_bert)

    def forward(self, txt_embeds, txt_masks, obj_embeds, obj_locs, obj_masks, output_attentions=False,
                output_hidden_states=False, **kwargs):
        pairwise_locs = calc_pairwise_locs(obj_locs[:, :, :3], obj_locs[:, :, 3:], pairwise_rel_type=self.pairwise_rel_type)
        out = obj_embeds
        loc = self.loc_layers[0](obj_locs).to(obj_embeds.dtype)   # identical every layer: computed once
        for layer in self.layers:
            out = out + loc
            out, _, _ = layer(out, txt_embeds, pairwise_locs, tgt_key_padding_mask=obj_masks.logical_not(),
                              memory_key_padding_mask=txt_masks.logical_not())
        return txt_embeds, out
