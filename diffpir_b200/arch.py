"""UNet architecture description shared by the host adapter and the weight packer.

Mirrors how the reference builds its denoiser:
  main_ddpir.py:219-230 (the two model configs), utils/utils_model.py:353-383 (argparse defaults),
  guided_diffusion/script_util.py:130-184 (create_model: channel_mult by image size, attention_ds,
  learn_sigma -> 6 output channels), guided_diffusion/unet.py:396-618 (UNetModel.__init__).
The C++ engine (csrc/engine.cu) replays the same constructor loop from the same fields, and
`param_specs` lists every checkpoint tensor (name, shape) in state_dict order.
"""
from dataclasses import dataclass, field
from typing import List, Tuple


@dataclass
class UNetConfig:
    image_size: int = 256
    in_channels: int = 3
    model_channels: int = 128
    out_channels: int = 6
    num_res_blocks: int = 1
    attention_resolutions: str = "16"
    channel_mult: Tuple[int, ...] = ()
    num_heads: int = 4
    num_head_channels: int = 64
    num_heads_upsample: int = -1
    use_scale_shift_norm: bool = True
    resblock_updown: bool = True
    use_new_attention_order: bool = False

    def __post_init__(self):
        if not self.channel_mult:
            # script_util.py:147-158
            table = {512: (0.5, 1, 1, 2, 2, 4, 4), 256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4)}
            if self.image_size not in table:
                raise ValueError(f"unsupported image size: {self.image_size}")
            self.channel_mult = table[self.image_size]
        elif isinstance(self.channel_mult, str):
            self.channel_mult = tuple(int(c) for c in self.channel_mult.split(","))
        self.channel_mult = tuple(self.channel_mult)

    @property
    def attention_ds(self):
        # script_util.py:161-164
        return tuple(self.image_size // int(r) for r in str(self.attention_resolutions).split(","))

    @property
    def time_embed_dim(self):
        return self.model_channels * 4

    def as_dict(self):
        return dict(image_size=self.image_size, in_channels=self.in_channels, model_channels=self.model_channels,
                    out_channels=self.out_channels, num_res_blocks=self.num_res_blocks,
                    attention_resolutions=self.attention_resolutions, channel_mult=tuple(self.channel_mult),
                    num_heads=self.num_heads, num_head_channels=self.num_head_channels,
                    num_heads_upsample=self.num_heads_upsample, use_scale_shift_norm=self.use_scale_shift_norm,
                    resblock_updown=self.resblock_updown, use_new_attention_order=self.use_new_attention_order)


def ffhq_config():
    """diffusion_ffhq_10m (main_ddpir.py:219-224)."""
    return UNetConfig(model_channels=128, num_res_blocks=1, attention_resolutions="16")


def imagenet_config():
    """256x256_diffusion_uncond (main_ddpir.py:225-230)."""
    return UNetConfig(model_channels=256, num_res_blocks=2, attention_resolutions="8,16,32")


def model_config_by_name(name):
    return ffhq_config() if name == "diffusion_ffhq_10m" else imagenet_config()


@dataclass
class Layer:
    kind: str            # "conv" | "res" | "attn"
    prefix: str          # state_dict prefix, e.g. "input_blocks.3.0."
    cin: int = 0
    cout: int = 0
    updown: str = ""     # "", "up", "down"
    heads: int = 0


def build_layers(cfg: UNetConfig):
    """Replays UNetModel.__init__ (unet.py:479-610): returns (inputs, middle, outputs), each a list of
    top-level blocks, each block a list of Layer."""
    mc, cm, nrb = cfg.model_channels, cfg.channel_mult, cfg.num_res_blocks
    att = cfg.attention_ds
    nh_up = cfg.num_heads if cfg.num_heads_upsample == -1 else cfg.num_heads_upsample
    if not (cfg.resblock_updown and cfg.use_scale_shift_norm):
        raise NotImplementedError("DiffPIR models use resblock_updown=True, use_scale_shift_norm=True")

    def heads(ch, nh):
        if cfg.num_head_channels == -1:
            return nh
        if ch % cfg.num_head_channels:
            raise ValueError("channels not divisible by num_head_channels")
        return ch // cfg.num_head_channels

    ch = int(cm[0] * mc)
    inputs = [[Layer("conv", "input_blocks.0.0.", cfg.in_channels, ch)]]
    chans = [ch]
    ds = 1
    for level, mult in enumerate(cm):
        for _ in range(nrb):
            i = len(inputs)
            block = [Layer("res", f"input_blocks.{i}.0.", ch, int(mult * mc))]
            ch = int(mult * mc)
            if ds in att:
                block.append(Layer("attn", f"input_blocks.{i}.1.", ch, ch, heads=heads(ch, cfg.num_heads)))
            inputs.append(block)
            chans.append(ch)
        if level != len(cm) - 1:
            i = len(inputs)
            inputs.append([Layer("res", f"input_blocks.{i}.0.", ch, ch, updown="down")])
            chans.append(ch)
            ds *= 2
    middle = [[Layer("res", "middle_block.0.", ch, ch),
               Layer("attn", "middle_block.1.", ch, ch, heads=heads(ch, cfg.num_heads)),
               Layer("res", "middle_block.2.", ch, ch)]]
    outputs = []
    for level, mult in list(enumerate(cm))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            o = len(outputs)
            block = [Layer("res", f"output_blocks.{o}.0.", ch + ich, int(mc * mult))]
            ch = int(mc * mult)
            if ds in att:
                block.append(Layer("attn", f"output_blocks.{o}.{len(block)}.", ch, ch, heads=heads(ch, nh_up)))
            if level and i == nrb:
                block.append(Layer("res", f"output_blocks.{o}.{len(block)}.", ch, ch, updown="up"))
                ds //= 2
            outputs.append(block)
    return inputs, middle, outputs


def param_specs(cfg: UNetConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """Every checkpoint tensor (name, shape) in the order of the reference state_dict."""
    ted = cfg.time_embed_dim
    specs = [("time_embed.0.weight", (ted, cfg.model_channels)), ("time_embed.0.bias", (ted,)),
             ("time_embed.2.weight", (ted, ted)), ("time_embed.2.bias", (ted,))]

    def layer_specs(l: Layer):
        p = l.prefix
        if l.kind == "conv":
            return [(p + "weight", (l.cout, l.cin, 3, 3)), (p + "bias", (l.cout,))]
        if l.kind == "res":
            s = [(p + "in_layers.0.weight", (l.cin,)), (p + "in_layers.0.bias", (l.cin,)),
                 (p + "in_layers.2.weight", (l.cout, l.cin, 3, 3)), (p + "in_layers.2.bias", (l.cout,)),
                 (p + "emb_layers.1.weight", (2 * l.cout, ted)), (p + "emb_layers.1.bias", (2 * l.cout,)),
                 (p + "out_layers.0.weight", (l.cout,)), (p + "out_layers.0.bias", (l.cout,)),
                 (p + "out_layers.3.weight", (l.cout, l.cout, 3, 3)), (p + "out_layers.3.bias", (l.cout,))]
            if l.cin != l.cout:
                s += [(p + "skip_connection.weight", (l.cout, l.cin, 1, 1)), (p + "skip_connection.bias", (l.cout,))]
            return s
        return [(p + "norm.weight", (l.cin,)), (p + "norm.bias", (l.cin,)),
                (p + "qkv.weight", (3 * l.cin, l.cin, 1)), (p + "qkv.bias", (3 * l.cin,)),
                (p + "proj_out.weight", (l.cin, l.cin, 1)), (p + "proj_out.bias", (l.cin,))]

    inputs, middle, outputs = build_layers(cfg)
    for blocks in (inputs, middle, outputs):
        for block in blocks:
            for l in block:
                specs += layer_specs(l)
    ch0 = int(cfg.channel_mult[0] * cfg.model_channels)
    specs += [("out.0.weight", (ch0,)), ("out.0.bias", (ch0,)),
              ("out.2.weight", (cfg.out_channels, ch0, 3, 3)), ("out.2.bias", (cfg.out_channels,))]
    return specs


def flops_per_image(cfg: UNetConfig, hw: int = None) -> float:
    """Algorithmic FLOPs (2*MAC) of one UNet forward for one image (conv3x3 + conv1x1 + qkv/proj +
    attention matmuls + linears); reproduces SURVEY.md section 3.3 (387.93 GF FFHQ, 2239.67 GF ImageNet)."""
    hw = hw or cfg.image_size
    inputs, middle, outputs = build_layers(cfg)
    ted = cfg.time_embed_dim
    total = 2.0 * (cfg.model_channels * ted + ted * ted)
    res = hw
    for blocks in (inputs, middle, outputs):
        for block in blocks:
            for l in block:
                if l.kind == "conv":
                    total += 2.0 * res * res * 9 * l.cin * l.cout
                elif l.kind == "res":
                    if l.updown == "down":
                        res //= 2
                    elif l.updown == "up":
                        res *= 2
                    total += 2.0 * res * res * 9 * (l.cin * l.cout + l.cout * l.cout)
                    total += 2.0 * ted * 2 * l.cout
                    if l.cin != l.cout:
                        total += 2.0 * res * res * l.cin * l.cout
                else:
                    T = res * res
                    total += 2.0 * T * l.cin * 4 * l.cin + 2.0 * 2 * T * T * l.cin
    ch0 = int(cfg.channel_mult[0] * cfg.model_channels)
    total += 2.0 * hw * hw * 9 * ch0 * cfg.out_channels
    return total
