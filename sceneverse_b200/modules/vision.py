"""Object encoder of GPS (reference: modules/vision/pcd_openvocab_encoder.py:16-184)."""
import glob
import os

import torch
import torch.nn.functional as F
from torch import nn

from .layers import LayerNorm, TransformerSpatialEncoderLayer, calc_pairwise_locs, init_weights_bert, layer_repeat
from .pointnet import GPS_SPEC, PointNetPP
from .registry import VISION_REGISTRY


@VISION_REGISTRY.register()
class PointOpenVocabEncoder(nn.Module):
    def __init__(self, cfg, backbone='pointnet++', hidden_size=768, path=None, freeze=False, dim_feedforward=2048,
                 num_attention_heads=12, spatial_dim=5, num_layers=4, dim_loc=6, pairwise_rel_type='center',
                 use_matmul_label=False, mixup_strategy=None, mixup_stage1=None, mixup_stage2=None,
                 lang_type='bert', lang_path=None, attn_type='spatial', text_features=None):
        super().__init__()
        assert backbone in ['pointnet++']
        self.point_feature_extractor = PointNetPP(**GPS_SPEC)
        # open-vocabulary class head: 607 frozen text embeddings (pcd_openvocab_encoder.py:46-48)
        if text_features is None:
            vocab = f"scannet_607_{'bert-base-uncased' if lang_type == 'bert' else 'clip-ViT-B16'}_id.pth"
            text_features = torch.load(os.path.join(lang_path, vocab))
        self.register_buffer("text_features", text_features)
        self.dropout = nn.Dropout(0.1)
        self.attn_type = attn_type
        self.freeze = freeze
        if freeze:  # only what exists at this point is frozen, as in the reference (:53-56)
            for p in self.parameters():
                p.requires_grad = False
        self.sem_cls_embed_layer = nn.Sequential(nn.Linear(hidden_size, hidden_size), LayerNorm(hidden_size),
                                                 nn.Dropout(0.1))
        self.use_matmul_label = use_matmul_label
        self.sem_mask_embeddings = nn.Embedding(1, 768)
        if self.attn_type == 'spatial':
            layer = TransformerSpatialEncoderLayer(hidden_size, num_attention_heads, dim_feedforward=dim_feedforward,
                                                   dropout=0.1, activation='gelu', spatial_dim=spatial_dim,
                                                   spatial_multihead=True, spatial_attn_fusion='cond')
            self.spatial_encoder = layer_repeat(layer, num_layers)
            self.loc_layers = layer_repeat(nn.Sequential(nn.Linear(dim_loc, hidden_size), LayerNorm(hidden_size)), 1)
            self.pairwise_rel_type = pairwise_rel_type
            self.spatial_dim = spatial_dim
        self.apply(init_weights_bert)
        if path is not None:
            ckpts = glob.glob(os.path.join(path, '*.bin'))
            if ckpts:
                for ckpt in ckpts:
                    self.load_state_dict(torch.load(ckpt, map_location='cpu'), strict=False)
            elif path.endswith('.pth'):
                self.load_state_dict(torch.load(path), strict=False)

    def point_cls_head(self, x):
        return x @ self.text_features.t()

    def freeze_bn(self, m):
        for layer in m.modules():
            if isinstance(layer, nn.BatchNorm2d):
                layer.eval()

    def forward(self, obj_pcds, obj_locs, obj_masks, obj_sem_masks, obj_labels=None, cur_step=None, max_steps=None,
                **kwargs):
        if self.freeze:
            self.freeze_bn(self.point_feature_extractor)
        B, O = obj_pcds.shape[:2]
        obj_embeds = self.point_feature_extractor(obj_pcds.reshape(B * O, *obj_pcds.shape[2:])).view(B, O, -1)
        obj_embeds = self.dropout(obj_embeds)
        if self.freeze:
            obj_embeds = obj_embeds.detach()
        obj_sem_cls = F.softmax(self.point_cls_head(obj_embeds), dim=2).detach()
        obj_embeds_pre = obj_embeds
        if self.attn_type == 'spatial':
            pairwise_locs = calc_pairwise_locs(obj_locs[:, :, :3], obj_locs[:, :, 3:],
                                               pairwise_rel_type=self.pairwise_rel_type, spatial_dist_norm=True,
                                               spatial_dim=self.spatial_dim)
            key_padding = obj_masks.logical_not()
            # the reference re-evaluates loc_layers[0](obj_locs) before every layer; same module, same input, no
            # dropout inside -> the same tensor each time, so it is computed once (autograd sums the per-layer uses)
            loc = self.loc_layers[0](obj_locs)
            # bf16 mode: the fp32 PointNet++ features enter the token stream in the compute dtype here, once — an fp32
            # residual stream would be re-cast to bf16 in front of every GEMM / LayerNorm of every layer, and back
            obj_embeds = obj_embeds.to(loc.dtype)
            for layer in self.spatial_encoder:
                obj_embeds = obj_embeds + loc
                obj_embeds, _ = layer(obj_embeds, pairwise_locs, tgt_key_padding_mask=key_padding)
        return obj_embeds, obj_embeds_pre, obj_sem_cls
