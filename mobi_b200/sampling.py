"""Shared machinery of the DDIM / PLMS drop-in samplers: schedule buffers, the fused CUDA step
(input assembly + CFG combine + x0 / x_prev update in two tiny kernels) and an optional CUDA graph of the
UNet call, so the 50-step loop issues no Python-side tensor math at all.
"""
import numpy as np
import torch

from . import ops


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True):
    """The DDPM timesteps a DDIM run visits (same values as ldm/modules/diffusionmodules/util.py:46-60): "uniform" takes
    every (T // S)-th step, "quad" spaces them quadratically over the first 80 % of the schedule; both are shifted by one so
    that the final alpha of a run is the first alpha of the DDPM schedule."""
    if ddim_discr_method == "uniform":
        stride = num_ddpm_timesteps // num_ddim_timesteps
        steps = np.arange(0, num_ddpm_timesteps, stride)
    elif ddim_discr_method == "quad":
        steps = np.square(np.linspace(0, np.sqrt(num_ddpm_timesteps * .8), num_ddim_timesteps)).astype(int)
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    steps = np.asarray(steps) + 1
    if verbose:
        print("DDIM timesteps: %s" % (steps,))
    return steps


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=True):
    """(sigmas, alphas, alphas_prev) of a DDIM run, with the reference's dtypes (util.py:63-74): alphas is the float32
    gather of the cumulative products, alphas_prev and sigmas come out of float64 numpy arithmetic; the alpha before the
    first visited step is alphacums[0], not 1."""
    alphacums = np.asarray(alphacums, dtype=np.float32)
    alphas = alphacums[ddim_timesteps]
    earlier = alphacums[ddim_timesteps[:-1]]
    # the reference builds this array from a Python list of floats, i.e. float64 - except for a one-step run, where the
    # list holds a single float32 scalar and numpy keeps that dtype
    alphas_prev = np.concatenate([alphacums[:1], earlier]).astype(np.float64 if earlier.size else np.float32)
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    if verbose:
        print("DDIM schedule: eta %s, a_t %s, a_(t-1) %s, sigma_t %s" % (eta, alphas, alphas_prev, sigmas))
    return sigmas, alphas, alphas_prev


class StepCoefficients:
    """The four per-step scalars of ddim.py:195-212, evaluated in float32 like the reference's 0-d tensors."""

    def __init__(self, alphas, alphas_prev, sqrt_one_minus_alphas, sigmas, index):
        f = np.float32
        a_t, a_prev, sigma_t = f(alphas[index]), f(alphas_prev[index]), f(sigmas[index])
        self.sqrt_one_minus_at = float(f(sqrt_one_minus_alphas[index]))
        self.sqrt_at = float(np.sqrt(a_t))
        self.sqrt_a_prev = float(np.sqrt(a_prev))
        self.dir_coef = float(np.sqrt(f(1.0) - a_prev - sigma_t * sigma_t))
        self.sigma = float(sigma_t)


class SamplerBase(object):
    def __init__(self, model, schedule="linear", use_cuda_graph=None, **kwargs):
        super().__init__()
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        self.use_cuda_graph = use_cuda_graph
        self._graph = self._graph_ctx = None
        self._static_key = None
        self.launches = 0  # UNet evaluations issued (for bench accounting)

    def register_buffer(self, name, attr):
        """The reference forces .to("cuda") here (ddim.py:19-23); the drop-in follows the model's device."""
        if type(attr) == torch.Tensor and attr.device != self.model.device:
            attr = attr.to(self.model.device)
        setattr(self, name, attr)

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        """Builds every schedule attribute the reference samplers expose (ddim.py:25-54 / plms.py:24-55) under the same
        names.  Device buffers are only what callers of the reference read back; the step kernels take host scalars, so
        the DDIM arrays stay numpy and nothing is indexed on the device during the loop."""
        model = self.model
        if model.alphas_cumprod.shape[0] != self.ddpm_num_timesteps:
            raise AssertionError("alphas have to be defined for each timestep")
        self.ddim_timesteps = make_ddim_timesteps(ddim_discretize, ddim_num_steps, self.ddpm_num_timesteps, verbose=verbose)
        ac = model.alphas_cumprod.detach().cpu()
        acp = model.alphas_cumprod_prev.detach().cpu()
        on_device = lambda t: torch.as_tensor(t).clone().detach().to(dtype=torch.float32, device=model.device)   # noqa: E731
        derived = {
            "betas": model.betas, "alphas_cumprod": model.alphas_cumprod, "alphas_cumprod_prev": model.alphas_cumprod_prev,
            "sqrt_alphas_cumprod": np.sqrt(ac), "sqrt_one_minus_alphas_cumprod": np.sqrt(1. - ac),
            "log_one_minus_alphas_cumprod": np.log(1. - ac), "sqrt_recip_alphas_cumprod": np.sqrt(1. / ac),
            "sqrt_recipm1_alphas_cumprod": np.sqrt(1. / ac - 1),
        }
        for name, value in derived.items():
            self.register_buffer(name, on_device(value))
        self.ddim_sigmas, self.ddim_alphas, self.ddim_alphas_prev = make_ddim_sampling_parameters(
            alphacums=ac.numpy(), ddim_timesteps=self.ddim_timesteps, eta=ddim_eta, verbose=verbose)
        self.ddim_sqrt_one_minus_alphas = np.sqrt(np.float32(1.) - self.ddim_alphas)
        self._sqrt_ac_host = np.sqrt(ac.numpy())
        self._sqrt_1mac_host = np.sqrt(np.float32(1.) - ac.numpy())
        sigmas_full = ddim_eta * torch.sqrt((1 - acp) / (1 - ac) * (1 - ac / acp))
        self.register_buffer("ddim_sigmas_for_original_num_steps", sigmas_full.to(model.device))

    # ------------------------------------------------------------------ model evaluation
    def _native_unet(self):
        from .openaimodel import UNetModel
        inner = getattr(getattr(self.model, "model", None), "diffusion_model", None)
        return inner if isinstance(inner, UNetModel) else None

    def _setup_eval(self, b, shape_rest, cond, uc, scale):
        """Static buffers for the sampling runs of one shape and, for the native UNet, two CUDA graphs that are captured
        ONCE per (shape, CFG) and reused by every later sample() call of this sampler:
          graph_ctx   context-only work (attn2 vectors, adapter tables) from the static conditioning buffer;
          graph_unet  one apply_model evaluation reading the static x_in / t_in buffers and those tables."""
        dev = self.model.device
        self._cfg = not (uc is None or scale == 1.)
        rows = 2 * b if self._cfg else b
        if isinstance(cond, dict) or isinstance(cond, list):
            raise NotImplementedError("mobi_b200 samplers take the conditioning as one [B, n, ctx] tensor")
        c_new = torch.cat([uc, cond]) if self._cfg else cond
        c_new = c_new.to(dev).float()
        unet = self._native_unet()
        want_graph = self.use_cuda_graph if self.use_cuda_graph is not None else (unet is not None)
        use_graph = bool(want_graph and unet is not None and dev.type == "cuda")
        key = (rows, tuple(shape_rest), self._cfg, tuple(c_new.shape), use_graph, id(unet),
               getattr(unet, "_pack_serial", None))
        if getattr(self, "_static_key", None) == key:
            self._c_in.copy_(c_new)
            if self._graph_ctx is not None:
                self._graph_ctx.replay()
                unet.pin_context(self._c_in, self._ctx_tabs_static)
            elif unet is not None:
                unet.pin_context(self._c_in, unet.prepare_context(self._c_in))
            return
        self._static_key = None
        self._graph = self._graph_ctx = None
        self._eps = None
        self._x_in = torch.empty((rows,) + tuple(shape_rest), device=dev, dtype=torch.float32)
        self._t_in = torch.empty((rows,), device=dev, dtype=torch.long)
        self._c_in = c_new.contiguous().clone()
        if unet is not None:
            unet.pin_context(self._c_in, unet.prepare_context(self._c_in))
        if use_graph:
            self._t_in.fill_(1)
            self._x_in.zero_()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):  # warm-up: lazy packing, cudaFuncSetAttribute, allocator pools
                    self._apply_model(self._x_in, self._t_in, self._c_in)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g_ctx = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_ctx):
                self._ctx_tabs_static = unet.prepare_context(self._c_in)
            unet.pin_context(self._c_in, self._ctx_tabs_static)
            g = torch.cuda.CUDAGraph()
            before = ops.Stats.launches
            with torch.cuda.graph(g, pool=g_ctx.pool()):
                self._eps = self._apply_model(self._x_in, self._t_in, self._c_in)
            self._graph_kernels = ops.Stats.launches - before  # kernels replayed per UNet evaluation
            self._graph, self._graph_ctx = g, g_ctx
            g_ctx.replay()  # capture does not execute: fill the tables for this run
        self._static_key = key

    def _apply_model(self, x_in, t_in, c_in):
        """apply_model with the native UNet told that the two CFG halves of x_in / t_in are identical copies."""
        unet = self._native_unet()
        if unet is not None:
            unet.cfg_shared_halves = bool(self._cfg)
        try:
            return self.model.apply_model(x_in, t_in, c_in)
        finally:
            if unet is not None:
                unet.cfg_shared_halves = False

    def _eval_model(self, step):
        """eps for the current self._x_in at timestep `step` ([uncond ; cond] rows under CFG)."""
        self._t_in.fill_(int(step))
        self.launches += 1
        if self._graph is not None:
            self._graph.replay()
            ops.Stats.launches += self._graph_kernels
            return self._eps
        return self._apply_model(self._x_in, self._t_in, self._c_in).float().contiguous()

    @staticmethod
    def _rest_from_kwargs(kwargs):
        if "test_model_kwargs" in kwargs:
            kw = kwargs["test_model_kwargs"]
            return kw["inpaint_image"].float().contiguous(), kw["inpaint_mask"].float().contiguous()
        if "rest" in kwargs:
            return kwargs["rest"].float().contiguous(), None
        if "inpaint_image" in kwargs:
            return kwargs["inpaint_image"].float().contiguous(), kwargs["inpaint_mask"].float().contiguous()
        raise Exception("kwargs must contain either 'test_model_kwargs' or 'rest' key")

    @staticmethod
    def _check_unsupported(quantize_denoised, score_corrector, noise_dropout):
        if quantize_denoised or score_corrector is not None or noise_dropout > 0.:
            raise NotImplementedError(
                "mobi_b200 samplers: quantize_denoised / score_corrector / noise_dropout are not used by MObI")
