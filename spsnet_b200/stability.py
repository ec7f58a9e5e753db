"""SPSNet stability generator, inference path: per-point stability `stds` from the raw cloud.

Mirrors reference stability_generate/model.py: `Surface_PW_feature` (:34-168, one `PointnetSampling` set-abstraction
layer with identity sampling, i.e. every point is a centre: M = N = 16384, radii 0.2 / 0.8, nsample 16 / 32 -- the most
expensive ball-query shape of the whole model), `Encoder_surface_feature` (:171-184, two Linear(64 -> 8) heads) and the
eval branch of `Generate_center.forward` (:545-588):

    soc_feature = SA(points)                               (B, N, 64)
    logvar      = fc2(soc_feature)                         (B, N, 8)
    stds        = sum_k exp(0.5 * logvar_k)                (B, N)       -> batch_dict['stds']

`stds` feeds SPSNet-IA's stability-aware top-k sampling (`PAGNet_Backbone`, pointnet2_modules.py:293-305).  Same
constructor arguments, `state_dict` layout (`feature_extract.SA_modules.*`, `feature_encoder.fc{1,2}.*`,
`obj_encoder.*`) and `batch_dict` keys as the reference, so generator checkpoints load unchanged; the training branch
(losses, target assignment, reparametrisation) is out of scope and raises.

Fused inference: the SA layer is the same tensor-core path as the backbone's layer 0 (split-fp16, fp32-grade), and the
`fc2` head runs as one point-wise tensor-core GEMM on the layer's point-major fp16 [values | residuals] rows.
"""
from __future__ import annotations

import copy
from typing import Any

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointnet2_modules as pm
from . import pointnet2_utils as pu
from .backbone import Cfg

# reference: stability_generate/cfgs/sf_unc.yaml:52-85
SF_UNC_MODEL = {
    "SF_FEATURE_DIM": 64,
    "LATENT_DIM": 8,
    "SA_CONFIG": {
        "NPOINT_LIST": [[16384]],
        "SAMPLE_RANGE_LIST": [[-1]],
        "SAMPLE_METHOD_LIST": [["D-FPS"]],
        "RADIUS_LIST": [[0.2, 0.8]],
        "NSAMPLE_LIST": [[16, 32]],
        "MLPS": [[[16, 16, 32], [32, 32, 64]]],
        "LAYER_TYPE": ["SA_Layer"],
        "DILATED_GROUP": [False],
        "AGGREGATION_MLPS": [[64]],
        "CONFIDENCE_MLPS": [[]],
        "LAYER_INPUT": [0],
        "CTR_INDEX": [-1],
        "MAX_TRANSLATE_RANGE": [3.0, 3.0, 2.0],
    },
    "GENERATOR": {"LATENT_DIM": 8, "PW_FEATURE_DIM": 64},
}


def sf_unc_cfg() -> Cfg:
    return Cfg(copy.deepcopy(SF_UNC_MODEL))


class Surface_PW_feature(nn.Module):
    """Point-wise surface feature extractor (reference stability_generate/model.py:34-168)."""

    def __init__(self, model_cfg: Any, input_channels: int = 4, **kwargs):
        super().__init__()
        if isinstance(model_cfg, dict) and not isinstance(model_cfg, Cfg):
            model_cfg = Cfg(model_cfg)
        self.model_cfg = model_cfg
        sa = model_cfg
        self.layer_types = sa.LAYER_TYPE
        self.ctr_idx_list = sa.CTR_INDEX
        self.layer_inputs = sa.LAYER_INPUT
        self.aggregation_mlps = sa.get("AGGREGATION_MLPS", None)
        self.SA_modules = nn.ModuleList()
        channel_in = input_channels - 3
        channel_out_list = [channel_in]
        for k in range(len(sa.NSAMPLE_LIST)):
            src = self.layer_inputs[k][-1] if isinstance(self.layer_inputs[k], list) else self.layer_inputs[k]
            channel_in = channel_out_list[src]
            mlps = [[channel_in] + list(spec) for spec in sa.MLPS[k]]
            channel_out = sum(spec[-1] for spec in mlps)
            agg = list(self.aggregation_mlps[k]) if self.aggregation_mlps and self.aggregation_mlps[k] else None
            if agg:
                channel_out = agg[-1]
            self.SA_modules.append(pm.PointnetSampling(
                npoint_list=sa.NPOINT_LIST[k], sample_range_list=sa.SAMPLE_RANGE_LIST[k],
                sample_type_list=sa.SAMPLE_METHOD_LIST[k], radii=sa.RADIUS_LIST[k], nsamples=sa.NSAMPLE_LIST[k],
                mlps=mlps, use_xyz=True, dilated_group=sa.DILATED_GROUP[k], aggregation_mlp=agg))
            channel_out_list.append(channel_out)

    def forward(self, batch_dict):
        batch_size = batch_dict["batch_size"]
        points = batch_dict["points"]
        if points.shape[0] % batch_size != 0:
            raise RuntimeError("every scene must hold the same number of points (reference asserts min == max)")
        batch_idx = points[:, 0]
        xyz = points[:, 1:4].contiguous().view(batch_size, -1, 3)
        features = points[:, 4:].contiguous() if points.size(-1) > 4 else None
        if features is not None:
            features = features.view(batch_size, -1, features.shape[-1]).permute(0, 2, 1).contiguous()
        bidx2d = batch_idx.view(batch_size, -1)
        encoder_xyz, encoder_features = [xyz], [features]
        encoder_coords = [torch.cat([bidx2d.unsqueeze(-1), xyz], dim=-1)]
        sa_ins_preds = []
        for i, module in enumerate(self.SA_modules):
            ctr_xyz = encoder_xyz[self.ctr_idx_list[i]] if self.ctr_idx_list[i] != -1 else None
            li_xyz, li_features, _ = module(encoder_xyz[self.layer_inputs[i]], encoder_features[self.layer_inputs[i]], None, ctr_xyz=ctr_xyz)
            encoder_xyz.append(li_xyz)
            encoder_coords.append(torch.cat([bidx2d[:, :li_xyz.shape[1], None].float(), li_xyz.view(batch_size, -1, 3)], dim=-1))
            encoder_features.append(li_features)
            sa_ins_preds.append([])
        batch_dict["encoder_xyz"] = encoder_xyz
        batch_dict["encoder_coords"] = encoder_coords
        batch_dict["sa_ins_preds"] = sa_ins_preds
        batch_dict["_soc_feature_cm"] = encoder_features[-1]                      # (B, C, N) + its fp16 rows
        batch_dict["soc_feature"] = encoder_features[-1].permute(0, 2, 1)         # (B, N, C), the reference's key
        return batch_dict


class Encoder_surface_feature(nn.Module):
    """mu / logvar heads (reference stability_generate/model.py:171-184)."""

    def __init__(self, input_channels: int, latent_size: int = 3):
        super().__init__()
        self.fc1 = nn.Linear(input_channels, latent_size)
        self.fc2 = nn.Linear(input_channels, latent_size)

    def forward(self, features):
        mu = self.fc1(features)
        logvar = self.fc2(features)
        dist = torch.distributions.Independent(torch.distributions.Normal(loc=mu, scale=torch.exp(logvar) + 3e-22), 1)
        return dist, mu, logvar

    def logvar_fused(self, features_cm: torch.Tensor) -> torch.Tensor:
        """logvar (B, N, latent) from the channel-major (B, C, N) features on the tensor-core point-wise GEMM."""
        B, _, N = features_cm.shape
        cache = self.__dict__.setdefault("_pw", {})
        key = (self.fc2.weight.data_ptr(), self.fc2.weight._version, self.fc2.bias._version)
        tw, lo = pm._get_twin(features_cm, 16)
        split = lo > 0
        if cache.get("key") != (key, split):
            cache["layer"] = pu.PwLayer(self.fc2.weight.detach().float().t().contiguous(), self.fc2.bias.detach().float(), False, split=split)
            cache["key"] = (key, split)
        out = torch.empty((B, N, self.fc2.out_features), dtype=torch.float32, device=features_cm.device)
        pu.pw_mma_forward(tw.view(B * N, -1), cache["layer"], xlo=lo, out_pm=out)
        return out


class Object_feat_encoder(nn.Module):
    """Centre decoder, used by the training branch only; kept for `state_dict` compatibility (reference :187-222)."""

    def __init__(self, model_cfg):
        super().__init__()
        latent_dim, fe = model_cfg.LATENT_DIM, model_cfg.PW_FEATURE_DIM
        w = int(256 * 0.25)
        self.fc1 = nn.Linear(fe + latent_dim, w)
        self.fc2 = nn.Linear(w, w)
        self.fc_ce1 = nn.Linear(w, w)
        self.fc_ce2 = nn.Linear(w, 3, bias=False)

    def forward(self, x, z):
        x = F.relu(self.fc1(torch.cat([x, z], dim=-1)))
        feat = F.relu(self.fc2(x))
        return self.fc_ce2(F.relu(self.fc_ce1(feat)))


class Generate_center(nn.Module):
    """Stability generator (reference stability_generate/model.py:225-588), eval branch."""

    def __init__(self, model_cfg: Any, **kwargs):
        super().__init__()
        if isinstance(model_cfg, dict) and not isinstance(model_cfg, Cfg):
            model_cfg = Cfg(model_cfg)
        self.model_cfg = model_cfg
        self.feature_extract = Surface_PW_feature(model_cfg.SA_CONFIG)
        self.feature_encoder = Encoder_surface_feature(input_channels=model_cfg.SF_FEATURE_DIM, latent_size=model_cfg.LATENT_DIM)
        self.obj_encoder = Object_feat_encoder(model_cfg.GENERATOR)
        self.register_buffer("global_step", torch.LongTensor(1).zero_())

    def forward(self, batch_dict, **kwargs):
        if kwargs.get("training", None) is not None:
            self.training = kwargs["training"]
        if self.training:
            raise NotImplementedError("the generator's training branch (losses, target assignment) is out of scope; call .eval()")
        batch_dict = self.feature_extract(batch_dict)
        feats_cm = batch_dict.pop("_soc_feature_cm")
        if pm._fused_ok(self, feats_cm) and pm.mlp_backend() == "mma":
            logvar = self.feature_encoder.logvar_fused(feats_cm)
        else:
            _, _, logvar = self.feature_encoder(batch_dict["soc_feature"])
        batch_dict["stds"] = torch.sum(torch.exp(0.5 * logvar), dim=-1)          # (B, N)   reference :577
        return batch_dict

    def load_params_from_file_wo_logger(self, filename, to_cpu=False):
        """reference :618-633: load the shape-compatible entries of checkpoint['model_state']."""
        ckpt = torch.load(filename, map_location=torch.device("cpu") if to_cpu else None)
        disk = ckpt["model_state"]
        state = self.state_dict()
        state.update({k: v for k, v in disk.items() if k in state and state[k].shape == v.shape})
        self.load_state_dict(state)


def delete_unstable_points(batch_dict, delete_number: int = 500):
    """The 'stability' branch of the point-deletion ablation in the reference's `PAGNet_encoding.forward`
    (pcdet/models/backbones_2d/map_to_bev/PAGNet_encoding.py:33-69): per scene keep the background points and the
    (n_fg - delete_number) foreground points (`fake_labels > 0`) with the LARGEST stds.  (The reference's other
    branches draw `torch.randperm` and are not reproducible; scenes with fewer foreground points than `delete_number`
    are returned unchanged here.)"""
    B = batch_dict["batch_size"]
    points, stds, fg = batch_dict["points"], batch_dict["stds"], batch_dict["fake_labels"] > 0
    out = []
    for b in range(B):
        m = points[:, 0] == b
        pts, fgm = points[m], fg[m]
        n_fg = int(fgm.sum())
        if n_fg > delete_number:
            _, keep = torch.topk(stds[b][fgm], n_fg - delete_number)
            out.append(torch.cat([pts[~fgm], pts[fgm][keep]]))
        else:
            out.append(pts)
    batch_dict["points"] = torch.cat(out, dim=0)
    return batch_dict


class SPSNetIAFrontEnd(nn.Module):
    """SPSNet-IA inference front end on the new path: stability generator -> per-point stds -> PAGNet backbone with
    stability-aware sampling (reference: PAGNet_encoding.forward without its point-deletion ablation, then
    PAGNet_Backbone.forward).  One `forward(batch_dict)`; CUDA-graph capturable like the backbone alone."""

    def __init__(self, generator: Generate_center, backbone: nn.Module):
        super().__init__()
        self.generator = generator
        self.backbone = backbone

    def forward(self, batch_dict):
        batch_dict = self.generator(batch_dict)
        for k in ("encoder_xyz", "encoder_coords", "sa_ins_preds", "soc_feature"):
            batch_dict.pop(k, None)
        return self.backbone(batch_dict)
