"""Model configuration: the HF-style attribute bag the reference's model classes read
(configs/r2r_model.json + the attributes bolted on at pretrain_src/train_r2r.py:102-105)."""
from types import SimpleNamespace

R2R_DEFAULTS = dict(
    hidden_size=768, num_attention_heads=12, intermediate_size=3072, hidden_act="gelu",
    hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, layer_norm_eps=1e-12,
    vocab_size=30522, max_position_embeddings=512, type_vocab_size=2, initializer_range=0.02,
    num_l_layers=9, num_x_layers=4, num_pano_layers=2, num_hidden_layers=12,
    image_feat_size=768, angle_feat_size=4, obj_feat_size=0, obj_prob_size=0, max_action_steps=100,
    update_lang_bert=True, use_lang2visn_attn=True, graph_sprels=True, glocal_fuse=True,
    bev_dim=21, bev_res=0.5, feat_dropout=0.4, output_attentions=False, output_hidden_states=False,
    pretrain_tasks=["mlm", "sap", "masksem"], sem_pred_token="cattn",
)


def make_config(**overrides):
    """Returns a transformers.PretrainedConfig when transformers is importable (what the reference's
    BertPreTrainedModel subclasses need), else a SimpleNamespace with the same attributes."""
    vals = dict(R2R_DEFAULTS)
    vals.update(overrides)
    try:
        from transformers import PretrainedConfig
        cfg = PretrainedConfig()
        for k, v in vals.items():
            setattr(cfg, k, v)
        return cfg
    except Exception:  # pragma: no cover
        return SimpleNamespace(**vals)
