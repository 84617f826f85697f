"""Functional CPU restatement of the guided-diffusion UNet forward (oracle; test infra only).

Follows /root/reference/guided_diffusion/unet.py:
  UNetModel.__init__ :396-618 (block structure), UNetModel.forward :634-663,
  ResBlock._forward :236-256, AttentionBlock._forward :299-305,
  QKVAttentionLegacy.forward :337-354, QKVAttention.forward :371-389,
  Upsample :81-110 (nearest x2, no conv), Downsample :113-140 (avg-pool 2, no conv)
and /root/reference/guided_diffusion/nn.py:
  GroupNorm32 :17-19 / normalization :93-100 (32 groups, eps 1e-5),
  timestep_embedding :103-121.

The network is evaluated straight from a checkpoint-style ``state_dict`` (names as in
the reference checkpoints) so no nn.Module of the reference is needed at run time.
Runs in the dtype of ``x`` (fp32 or fp64).
"""
import math

import torch
import torch.nn.functional as F


def ffhq_config():
    """main_ddpir.py:219-224 + utils/utils_model.py:353-383 + script_util.py:149-170."""
    return dict(image_size=256, in_channels=3, model_channels=128, out_channels=6, num_res_blocks=1,
                attention_resolutions="16", channel_mult=(1, 1, 2, 2, 4, 4), num_heads=4,
                num_head_channels=64, num_heads_upsample=-1, use_scale_shift_norm=True,
                resblock_updown=True, use_new_attention_order=False)


def imagenet_config():
    """main_ddpir.py:225-230."""
    c = ffhq_config()
    c.update(model_channels=256, num_res_blocks=2, attention_resolutions="8,16,32")
    return c


def _attention_ds(cfg):
    # script_util.py:161-164
    return tuple(cfg["image_size"] // int(r) for r in str(cfg["attention_resolutions"]).split(","))


def block_program(cfg):
    """Replays UNetModel.__init__ (unet.py:479-610) and returns, per top-level block, the list of
    sub-layers as tuples:
      ("conv", cin, cout) | ("res", cin, cout, updown) | ("attn", ch, heads)
    for input_blocks, middle_block, output_blocks."""
    mc = cfg["model_channels"]
    cm = cfg["channel_mult"]
    nrb = cfg["num_res_blocks"]
    att = _attention_ds(cfg)
    nhc = cfg["num_head_channels"]
    nh_up = cfg["num_heads"] if cfg["num_heads_upsample"] == -1 else cfg["num_heads_upsample"]

    def heads(ch, nh):
        return nh if nhc == -1 else ch // nhc

    assert cfg["resblock_updown"], "reference DiffPIR models use resblock_updown=True"
    ch = int(cm[0] * mc)
    inputs = [[("conv", cfg["in_channels"], ch)]]
    chans = [ch]
    ds = 1
    for level, mult in enumerate(cm):
        for _ in range(nrb):
            layers = [("res", ch, int(mult * mc), None)]
            ch = int(mult * mc)
            if ds in att:
                layers.append(("attn", ch, heads(ch, cfg["num_heads"])))
            inputs.append(layers)
            chans.append(ch)
        if level != len(cm) - 1:
            inputs.append([("res", ch, ch, "down")])
            chans.append(ch)
            ds *= 2
    middle = [("res", ch, ch, None), ("attn", ch, heads(ch, cfg["num_heads"])), ("res", ch, ch, None)]
    outputs = []
    for level, mult in list(enumerate(cm))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, int(mc * mult), None)]
            ch = int(mc * mult)
            if ds in att:
                layers.append(("attn", ch, heads(ch, nh_up)))
            if level and i == nrb:
                layers.append(("res", ch, ch, "up"))
                ds //= 2
            outputs.append(layers)
    return inputs, middle, outputs


def timestep_embedding(t, dim, max_period=10000):
    """nn.py:103-121 (freqs and args are float32 there; kept float32 then cast)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _gn(x, w, b):
    return F.group_norm(x, 32, w, b, eps=1e-5)


def _silu(x):
    return x * torch.sigmoid(x)


def resblock(sd, p, x, emb, updown):
    """unet.py:236-256, use_scale_shift_norm=True branch."""
    g = lambda k: sd[p + k].to(x.dtype)
    h = _silu(_gn(x, g("in_layers.0.weight"), g("in_layers.0.bias")))
    if updown == "down":
        h = F.avg_pool2d(h, 2, 2)
        x = F.avg_pool2d(x, 2, 2)
    elif updown == "up":
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    h = F.conv2d(h, g("in_layers.2.weight"), g("in_layers.2.bias"), padding=1)
    eo = F.linear(_silu(emb), g("emb_layers.1.weight"), g("emb_layers.1.bias"))[..., None, None]
    scale, shift = torch.chunk(eo, 2, dim=1)
    h = _gn(h, g("out_layers.0.weight"), g("out_layers.0.bias")) * (1 + scale) + shift
    h = F.conv2d(_silu(h), g("out_layers.3.weight"), g("out_layers.3.bias"), padding=1)
    if (p + "skip_connection.weight") in sd:
        w = g("skip_connection.weight")
        x = F.conv2d(x, w, g("skip_connection.bias"), padding=w.shape[-1] // 2)
    return x + h


def attention(sd, p, x, n_heads, new_order=False):
    """unet.py:299-305 and QKVAttentionLegacy :337-354 (QKVAttention :371-389 when new_order)."""
    g = lambda k: sd[p + k].to(x.dtype)
    b, c, *sp = x.shape
    xf = x.reshape(b, c, -1)
    qkv = F.conv1d(_gn(xf, g("norm.weight"), g("norm.bias")), g("qkv.weight"), g("qkv.bias"))
    bs, width, length = qkv.shape
    ch = width // (3 * n_heads)
    scale = 1 / math.sqrt(math.sqrt(ch))
    if new_order:
        q, k, v = qkv.chunk(3, dim=1)
        q = q.reshape(bs * n_heads, ch, length)
        k = k.reshape(bs * n_heads, ch, length)
        v = v.reshape(bs * n_heads, ch, length)
    else:
        q, k, v = qkv.reshape(bs * n_heads, ch * 3, length).split(ch, dim=1)
    w = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    w = torch.softmax(w, dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(bs, -1, length)
    h = F.conv1d(a, g("proj_out.weight"), g("proj_out.bias"))
    return (xf + h).reshape(b, c, *sp)


def _run_layers(sd, prefix, layers, h, emb, cfg):
    for j, layer in enumerate(layers):
        p = f"{prefix}{j}."
        if layer[0] == "conv":
            h = F.conv2d(h, sd[p + "weight"].to(h.dtype), sd[p + "bias"].to(h.dtype), padding=1)
        elif layer[0] == "res":
            h = resblock(sd, p, h, emb, layer[3])
        else:
            h = attention(sd, p, h, layer[2], cfg["use_new_attention_order"])
    return h


def unet_forward(sd, cfg, x, t):
    """UNetModel.forward, unet.py:634-663.  x: (B,3,H,W); t: (B,) integer or float timesteps."""
    inputs, middle, outputs = block_program(cfg)
    dt = x.dtype
    emb = timestep_embedding(t, cfg["model_channels"]).to(dt)
    emb = F.linear(emb, sd["time_embed.0.weight"].to(dt), sd["time_embed.0.bias"].to(dt))
    emb = F.linear(_silu(emb), sd["time_embed.2.weight"].to(dt), sd["time_embed.2.bias"].to(dt))
    hs = []
    h = x
    for i, layers in enumerate(inputs):
        h = _run_layers(sd, f"input_blocks.{i}.", layers, h, emb, cfg)
        hs.append(h)
    h = _run_layers(sd, "middle_block.", middle, h, emb, cfg)
    for i, layers in enumerate(outputs):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_layers(sd, f"output_blocks.{i}.", layers, h, emb, cfg)
    h = _silu(_gn(h, sd["out.0.weight"].to(dt), sd["out.0.bias"].to(dt)))
    return F.conv2d(h, sd["out.2.weight"].to(dt), sd["out.2.bias"].to(dt), padding=1)
