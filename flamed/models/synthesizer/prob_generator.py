"""Flow-matching code-decoder denoiser.

Drop-in for the reference's flamed/models/synthesizer/prob_generator.py: same module tree and
parameter names (504-key checkpoint layout), same `ProbGenerator.sample` signature.  The cond
fold/down-sampler, the N-step Euler loop and every block inside it run in sm_100a kernels
(tcgen05 GEMMs + fused memory-bound kernels, one CUDA graph per (B, L, nfe)) behind
flm_cond_prepare / flm_denoiser_sample.
"""
import torch
import torch.nn as nn

from ._engine import EngineOwner


def _indexed(mods):
    m = nn.Module()
    for k, v in mods.items():
        m.add_module(str(k), v)
    return m


class _ConvNeXt(nn.Module):
    def __init__(self, c, k):
        super().__init__()
        self.conv_1 = nn.Conv1d(c, c, k, padding=k // 2, groups=c)
        self.ln_1 = nn.GroupNorm(c, c)
        self.conv_2 = nn.Conv1d(c, c, 1)
        self.conv_3 = nn.Conv1d(c, c, 1)


class _ResBlock(nn.Module):
    def __init__(self, c, k):
        super().__init__()
        self.adaLN_modulation = _indexed({1: nn.Linear(c, 6 * c)})
        self.ln_conv = nn.LayerNorm(c, eps=1e-6)
        self.conv_in = _ConvNeXt(c, k)
        self.ln_mlp = nn.LayerNorm(c, eps=1e-6)
        self.mlp = _indexed({0: nn.Linear(c, c), 2: nn.Linear(c, c)})


class _FinalLayer(nn.Module):
    def __init__(self, c, out_c, k):
        super().__init__()
        self.adaLN_modulation = _indexed({1: nn.Linear(c, 5 * c)})
        self.conv_in = _ConvNeXt(c, k)
        self.conv_out = nn.Conv1d(c, out_c, 3, padding=1)


class SimpleMLPAdaLN(nn.Module):
    """parameter holder (reference prob_generator.py:267-365)"""

    def __init__(self, in_channels, model_channels, out_channels, spk_dim, num_res_blocks, kernel):
        super().__init__()
        self.time_embed = nn.Module()
        self.time_embed.add_module("mlp", _indexed({0: nn.Linear(256, model_channels),
                                                    2: nn.Linear(model_channels, model_channels)}))
        self.cond_embed = nn.Linear(spk_dim, model_channels)
        self.proj_in = nn.Linear(in_channels, model_channels)
        self.res_blocks = nn.ModuleList(_ResBlock(model_channels, kernel) for _ in range(num_res_blocks))
        self.final_layer = _FinalLayer(model_channels, out_channels, kernel)


class _CondDownSampler(nn.Module):
    def __init__(self, cin, cout, n_stages):
        super().__init__()
        self.resblocks, self.downblocks = nn.ModuleList(), nn.ModuleList()
        for _ in range(n_stages):
            rb = nn.Module()
            rb.add_module("block", nn.Module())
            rb.block.add_module("block", _indexed({0: nn.Conv1d(cin, cin, 1), 1: nn.GroupNorm(8, cin)}))
            self.resblocks.append(rb)
            self.downblocks.append(_indexed({0: nn.Conv1d(cin, cin // 2, 1), 1: nn.GroupNorm(8, cin // 2)}))
            cin //= 2
        self.proj_out = _indexed({0: nn.Linear(cin, cout)})


class ProbGenerator(EngineOwner):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.target_dim = config["target_dim"]
        self.sigma_min = float(config["sigma_min"])
        self.quantizer_encoding = nn.Module()
        self.quantizer_encoding.add_module("quantizer_emb", nn.Embedding(config["n_quantizers"], config["cond_dim"]))
        self.cond_downsampling = _CondDownSampler(config["n_quantizers"] * config["cond_dim"], config["target_dim"],
                                                  config["downsampling_stages"])
        self.denoiser = SimpleMLPAdaLN(config["target_dim"], config["hidden_dim"], config["target_dim"],
                                       config["spk_dim"], config["n_layers"], config["convnext"]["kernel_size"])
        self.noise_device = "cpu"
        # True / False / "auto": one CUDA graph per (B, L, nfe) pays off when the kernels are short
        # (launch-bound, few frames); big batches are launched directly (kernels of ~100 us hide the launches)
        self.use_cuda_graph = "auto"
        self.graph_max_rows = 16384

    def _build_engine(self, ctx):
        from flamed_tts_b200.engines import DenoiserEngine
        return DenoiserEngine(ctx, self.state_dict(), self.config, precision=self.precision)

    def compute_loss(self, *a, **k):
        raise NotImplementedError("training is out of scope of the B200 inference hot path")

    @torch.inference_mode()
    def sample(self, cond, spk, mask, nfe=4, temperature=1.0):
        """cond (B,Q,L,cond_dim) prior embeddings, spk (B,spk_dim), mask (B,L,1) True = valid
        -> latents (B,target_dim,L): a transposed view, exactly as the reference returns it
        (prob_generator.py:434-446)."""
        eng = self.engine()
        c = eng.cond_prepare(cond, mask)
        b, l, d = c.shape
        ts = torch.linspace(0, 1, nfe + 1)
        seed = 0
        if self.noise_device == "philox":  # drawn inside the x0 kernel from a seed (documented map, flamed_b200.h)
            noise, seed = None, int(torch.randint(0, 2 ** 62, (1,)).item())
        else:
            ndev = c.device if self.noise_device == "cuda" else "cpu"
            noise = torch.randn((b, l, self.target_dim), device=ndev)
        graph = (b * l <= self.graph_max_rows) if self.use_cuda_graph == "auto" else bool(self.use_cuda_graph)
        x = eng.sample(c, spk, noise, ts, temperature, use_graph=graph, seed=seed)
        return x.transpose(1, -1)
