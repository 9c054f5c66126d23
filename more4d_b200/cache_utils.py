"""TeaCache for the B200 DiT path — same interface and decision rule as the reference's
MoRe4D/models/cache_utils.py:19-74 and the hooks in wan_transformer4d.py:1200-1270: skip the
block stack when the accumulated, polynomially rescaled relative-L1 change of the modulated
timestep embedding e0 stays under a threshold, re-using the previous step's block residual.
The state machine is host Python by design ("state machine stays Python")."""
from __future__ import annotations

import numpy as np
import torch


def get_teacache_coefficients(model_name: str):
    name = model_name.lower()
    table = [
        (("wan2.1-t2v-1.3b", "wan2.1-fun-1.3b", "wan2.1-fun-v1.1-1.3b"),
         [-5.21862437e+04, 9.23041404e+03, -5.28275948e+02, 1.36987616e+01, -4.99875664e-02]),
        (("wan2.1-t2v-14b",),
         [-3.03318725e+05, 4.90537029e+04, -2.65530556e+03, 5.87365115e+01, -3.15583525e-01]),
        (("wan2.1-i2v-14b-480p",),
         [2.57151496e+05, -3.54229917e+04, 1.40286849e+03, -1.35890334e+01, 1.32517977e-01]),
        (("wan2.1-i2v-14b-720p", "wan2.1-fun-14b", "wan2.2-fun", "wan2.2-i2v-a14b", "wan2.2-t2v-a14b",
          "wan2.2-ti2v-5b"),
         [8.10705460e+03, 2.13393892e+03, -3.72934672e+02, 1.66203073e+01, -4.17769401e-02]),
    ]
    for keys, coeff in table:
        if any(k in name for k in keys):
            return coeff
    print(f"The model {model_name} is not supported by TeaCache.")
    return None


class TeaCache:
    def __init__(self, coefficients, num_steps: int, rel_l1_thresh: float = 0.0,
                 num_skip_start_steps: int = 0, offload: bool = True):
        if num_steps < 1:
            raise ValueError(f"`num_steps` must be greater than 0 but is {num_steps}.")
        if rel_l1_thresh < 0:
            raise ValueError(f"`rel_l1_thresh` must be greater than or equal to 0 but is {rel_l1_thresh}.")
        if num_skip_start_steps < 0 or num_skip_start_steps > num_steps:
            raise ValueError("`num_skip_start_steps` must be in [0, num_steps]")
        self.coefficients = coefficients
        self.num_steps = num_steps
        self.rel_l1_thresh = rel_l1_thresh
        self.num_skip_start_steps = num_skip_start_steps
        self.offload = offload
        self.rescale_func = np.poly1d(self.coefficients)
        self.reset()

    @staticmethod
    def compute_rel_l1_distance(prev: torch.Tensor, cur: torch.Tensor) -> float:
        return _rel_l1(prev, cur)

    def reset(self):
        _reset(self)

    def decide(self, modulated_inp: torch.Tensor, cond_flag: bool) -> bool:
        return teacache_decide(self, modulated_inp, cond_flag)

    def step_done(self, cond_flag: bool):
        teacache_step_done(self, cond_flag)


def _rel_l1(prev: torch.Tensor, cur: torch.Tensor) -> float:
    return ((cur - prev).abs().mean() / prev.abs().mean()).cpu().item()


def _reset(tc) -> None:
    tc.cnt = 0
    tc.should_calc = True
    tc.accumulated_rel_l1_distance = 0
    tc.previous_modulated_input = None
    tc.previous_residual = None
    tc.previous_residual_cond = None
    tc.previous_residual_uncond = None


def teacache_decide(tc, modulated_inp: torch.Tensor, cond_flag: bool) -> bool:
    """The decision block of wan_transformer4d.py:1201-1220 as a free function over the FIELDS
    of a TeaCache (`cnt`, `num_skip_start_steps`, `accumulated_rel_l1_distance`, `rescale_func`,
    `rel_l1_thresh`, `previous_modulated_input`, `should_calc`): `tc` may be this module's class
    or the reference's own MoRe4D/models/cache_utils.py:19-74 instance, which the pipeline /
    infer.py create (`enable_teacache`, t4d:961-970) and `install()` hands over unchanged."""
    if not cond_flag:
        return tc.should_calc
    if tc.cnt < tc.num_skip_start_steps or tc.previous_modulated_input is None:
        should_calc = True            # (the reference would raise on a None previous input)
        tc.accumulated_rel_l1_distance = 0
    else:
        d = _rel_l1(tc.previous_modulated_input, modulated_inp)
        tc.accumulated_rel_l1_distance += tc.rescale_func(d)
        if tc.accumulated_rel_l1_distance < tc.rel_l1_thresh:
            should_calc = False
        else:
            should_calc = True
            tc.accumulated_rel_l1_distance = 0
    tc.previous_modulated_input = modulated_inp
    tc.should_calc = should_calc
    return should_calc


def teacache_step_done(tc, cond_flag: bool) -> None:
    """wan_transformer4d.py:1336-1339."""
    if cond_flag:
        tc.cnt += 1
        if tc.cnt == tc.num_steps:
            _reset(tc)
