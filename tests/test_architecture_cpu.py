"""Architecture pins that need no GPU (modules built on the meta device): parameter counts against published model sizes and against
the installed transformers, state-dict key / shape identity between native modules and their oracles (the checks available
for the from-spec UNet / VAE restatements while diffusers is not installable)."""
import torch

from oracle import unet_oracle as UO
from oracle import vae_oracle as VO


def _sig(m):
    return [(k, tuple(v.shape)) for k, v in m.state_dict().items()]


def test_unet_sd21_param_count_and_keys():
    from dreamllm_b200.unet import UNet2DConditionModel
    with torch.device("meta"):
        m, r = UNet2DConditionModel(), UO.UNet2DConditionModel()
    assert sum(p.numel() for p in m.parameters()) == 865_910_724            # stabilityai/stable-diffusion-2-1-base UNet (published size)
    assert _sig(m) == _sig(r) and len(m.state_dict()) == 686
    keys = set(m.state_dict())
    # spot checks of diffusers' naming
    for k in ("conv_in.weight", "time_embedding.linear_1.weight", "down_blocks.0.attentions.1.transformer_blocks.0.attn2.to_k.weight",
              "down_blocks.2.downsamplers.0.conv.weight", "mid_block.attentions.0.proj_in.weight", "up_blocks.3.resnets.2.conv_shortcut.weight",
              "up_blocks.1.attentions.2.transformer_blocks.0.ff.net.0.proj.weight", "conv_norm_out.weight", "conv_out.bias"):
        assert k in keys, k
    sd = m.state_dict()
    assert sd["down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k.weight"].shape == (320, 1024)     # cross_attention_dim 1024
    assert sd["mid_block.attentions.0.proj_in.weight"].shape == (1280, 1280)                                  # use_linear_projection
    assert sd["down_blocks.1.attentions.0.transformer_blocks.0.ff.net.0.proj.weight"].shape == (5120, 640)    # GEGLU: 2 * 4 * 640


def test_vae_param_counts_and_keys():
    from dreamllm_b200.vae import AutoencoderKLDecoder, AutoencoderKLEncoder
    with torch.device("meta"):
        e, d = AutoencoderKLEncoder(), AutoencoderKLDecoder()
        re_, rd = VO.AutoencoderKLEncoder(), VO.AutoencoderKLDecoder()
    ne, nd = sum(p.numel() for p in e.parameters()), sum(p.numel() for p in d.parameters())
    assert ne + nd == 83_653_863                                             # SD AutoencoderKL (published size)
    assert _sig(e) == _sig(re_) and _sig(d) == _sig(rd)
    assert all(k.startswith(("encoder.", "quant_conv.")) for k in e.state_dict())
    assert all(k.startswith(("decoder.", "post_quant_conv.")) for k in d.state_dict())
    assert e.state_dict()["encoder.mid_block.attentions.0.to_q.weight"].shape == (512, 512)


def test_clip_tower_matches_transformers_clip_vit_l14_336():
    from transformers import CLIPVisionConfig
    from transformers import CLIPVisionModel as HF

    from dreamllm_b200.clip_vision import CLIPVisionConfigLite, CLIPVisionModel
    kw = dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16, image_size=336, patch_size=14)
    with torch.device("meta"):
        ours = CLIPVisionModel(CLIPVisionConfigLite(**kw))
        hf = HF(CLIPVisionConfig(**kw))
    mine, theirs = dict(_sig(ours)), dict(_sig(hf))
    theirs.pop("vision_model.embeddings.position_ids", None)                  # non-persistent in recent transformers
    mine.pop("vision_model.embeddings.position_ids", None)
    assert mine == theirs
    assert sum(p.numel() for p in ours.parameters()) == sum(p.numel() for p in hf.parameters()) == 303_507_456   # openai/clip-vit-large-patch14-336 vision tower


def test_vicuna_7b_param_count_and_reference_keys():
    from dreamllm_b200.modeling_dreamllm import DreamLLMConfig, DreamLLMForCausalMLM
    with torch.device("meta"):
        m = DreamLLMForCausalMLM(DreamLLMConfig.vicuna_7b())
    n = sum(p.numel() for p in m.parameters())
    assert n == 6_738_415_616 + 8 * 4096 * 2                                  # LLaMA-7B (32 000 vocab) + 8 added tokens in embed + lm_head
    sd = m.state_dict()
    assert sd["model.layers.31.self_attn.rotary_emb.inv_freq"].shape == (64,) and sd["lm_head.weight"].shape == (32008, 4096)
