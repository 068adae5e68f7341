"""Device-side input pipeline: the `data_dict` of one training step built ON THE GPU from ragged raw scene
data, replacing the dataloader-worker loops of data/datasets/base.py:697-741 (`_obj_processing_post`),
data/datasets/dataset_wrapper.py:38-111 (padding, masks) and data/data_utils.py:76-121 (`random_word`, `random_point_cloud`).

Input (what a scan loader hands over after reading the files): every object's raw points concatenated (total, 6) [xyz rgb],
a CSR offset per object SLOT (B * max_obj_len + 1; padded slots are empty ranges), the tokenised captions.  Output: the keys
`OpenVocab.forward` consumes, same dtypes and padding conventions as the reference (padded objects = all-ones points, zero
locs, label -100, mask False).  Randomness is a counter hash of (seed, index): reproducible and order independent."""
import torch

from . import _lib


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def prepare_objects(raw_points, slot_offsets, B, O, P=1024, seed=0, return_indices=False):
    """raw_points (total,6) f32 cuda, slot_offsets (B*O+1) int64 cuda -> obj_fts (B,O,P,6) f32, obj_locs (B,O,6) f32,
    obj_masks (B,O) bool [, sample indices (B,O,P) i32]."""
    assert raw_points.is_cuda and raw_points.dtype == torch.float32 and raw_points.dim() == 2 and raw_points.shape[1] == 6
    assert slot_offsets.dtype == torch.int64 and slot_offsets.numel() == B * O + 1
    raw_points, slot_offsets = raw_points.contiguous(), slot_offsets.contiguous()
    dev = raw_points.device
    fts = torch.empty((B, O, P, 6), dtype=torch.float32, device=dev)
    locs = torch.empty((B, O, 6), dtype=torch.float32, device=dev)
    masks = torch.empty((B, O), dtype=torch.uint8, device=dev)
    idx = torch.empty((B, O, P), dtype=torch.int32, device=dev) if return_indices else None
    lib = _lib.pointops()
    with torch.cuda.device(dev):
        st = lib.sv_scene_prep_f32(raw_points.data_ptr(), slot_offsets.data_ptr(), B * O, P, int(seed) & 0x7FFFFFFFFFFFFFFF,
                                   fts.data_ptr(), locs.data_ptr(), masks.data_ptr(),
                                   idx.data_ptr() if idx is not None else None, _stream(raw_points))
    _lib.check(lib, st, "sv_scene_prep_f32")
    out = (fts, locs, masks.bool())
    return out + (idx,) if return_indices else out


def mask_tokens(txt_ids, txt_masks, mask_ratio=0.15, mask_token_id=103, vocab_size=30522, seed=0):
    """BERT masked-LM corruption: (B,L) int64 ids / masks -> (masked ids, labels with -1 = not supervised)."""
    assert txt_ids.is_cuda and txt_ids.dtype == torch.int64 and txt_masks.dtype == torch.int64 and txt_ids.shape == txt_masks.shape
    ids, am = txt_ids.contiguous(), txt_masks.contiguous()
    out, lab = torch.empty_like(ids), torch.empty_like(ids)
    lib = _lib.pointops()
    with torch.cuda.device(ids.device):
        st = lib.sv_token_mask(ids.data_ptr(), am.data_ptr(), ids.numel(), float(mask_ratio), int(mask_token_id), int(vocab_size),
                               int(seed) & 0x7FFFFFFFFFFFFFFF, out.data_ptr(), lab.data_ptr(), _stream(ids))
    _lib.check(lib, st, "sv_token_mask")
    return out, lab


def mask_objects(obj_masks, drop_ratio=0.1, seed=0):
    """random_point_cloud: (B,O) bool -> (B,O) bool, True = keep the object's semantic features."""
    v = obj_masks.to(torch.uint8).contiguous()
    out = torch.empty_like(v)
    lib = _lib.pointops()
    with torch.cuda.device(v.device):
        st = lib.sv_coin_mask(v.data_ptr(), v.numel(), float(drop_ratio), int(seed) & 0x7FFFFFFFFFFFFFFF, out.data_ptr(), _stream(v))
    _lib.check(lib, st, "sv_coin_mask")
    return out.bool()


def build_data_dict(raw_points, slot_offsets, obj_labels, txt_ids, txt_masks, tgt_object_id, B, O, P=1024, seed=0,
                    txt_mask_ratio=0.15, pc_mask_ratio=0.1, scene_txt_ids=None, scene_txt_masks=None):
    """Everything `OpenVocab.forward` + the losses read, in three launches.  obj_labels: (B,O) int64 with -100 on padded slots
    (dataset_wrapper.py:70-71); tgt_object_id (B,1) int64."""
    fts, locs, masks = prepare_objects(raw_points, slot_offsets, B, O, P, seed)
    ids, labels = mask_tokens(txt_ids, txt_masks, txt_mask_ratio, seed=seed + 1)
    d = {"obj_fts": fts, "obj_locs": locs, "obj_masks": masks, "obj_sem_masks": mask_objects(masks, pc_mask_ratio, seed + 2),
         "obj_labels": obj_labels, "txt_ids": ids, "txt_masks": txt_masks, "masked_lm_labels": labels, "tgt_object_id": tgt_object_id}
    if scene_txt_ids is not None:
        d["scene_txt_ids"], d["scene_txt_masks"] = scene_txt_ids, scene_txt_masks
    return d
