"""Host-side schedule logic of DYffusion: diffusion-step -> interpolation-time map and the sampling-schedule
string parser (reference: src/diffusion/dyffusion.py:44-138 and :245-333).  Pure Python, exact."""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Union

Number = Union[int, float]


class DiffusionSchedule:
    def __init__(self, timesteps: int, schedule: str, additional_interpolation_steps: int,
                 additional_interpolation_steps_factor: int, interpolate_before_t1: bool):
        horizon = timesteps
        assert horizon > 1, f"horizon must be > 1, but got {horizon}. Please use datamodule.horizon with > 1"
        self.kind = schedule
        self.aux_steps = 0       # k of 'before_t1_only'
        self.factor = 0          # factor of 'linear'
        self.offset = 0
        if schedule == "linear":
            assert additional_interpolation_steps == 0, \
                "additional_interpolation_steps must be 0 when using linear schedule"
            self.factor = additional_interpolation_steps_factor
            if interpolate_before_t1:
                n_interp = horizon - 1
            else:
                n_interp, self.offset = horizon - 2, additional_interpolation_steps_factor
            extra = additional_interpolation_steps_factor * n_interp
        elif schedule == "before_t1_only":
            assert additional_interpolation_steps_factor == 0, \
                "additional_interpolation_steps_factor must be 0 when using before_t1_only schedule"
            assert interpolate_before_t1, "interpolate_before_t1 must be True when using before_t1_only schedule"
            extra = self.aux_steps = additional_interpolation_steps
        else:
            raise ValueError(f"Invalid schedule: {schedule}")
        self.additional_diffusion_steps = extra
        self.num_timesteps = horizon + extra
        table = {d: self.interpolation_time(d) for d in range(1, self.num_timesteps)}
        self.dynamical_steps: Dict[int, Number] = {d: t for d, t in table.items() if float(t).is_integer()}
        self.artificial_interpolation_steps: Dict[int, Number] = {
            d: t for d, t in table.items() if not float(t).is_integer()}
        self.i_to_diffusion_step = {t: d for d, t in table.items()}

    def interpolation_time(self, d: Number) -> Number:
        """diffusion_step_to_interpolation_step for python scalars (:101-138)."""
        assert 0 <= d <= self.num_timesteps - 1, \
            f"diffusion_step must be in [1, num_timesteps-1]=[1, {self.num_timesteps - 1}], but got {d}"
        if self.kind == "linear":
            return (d + self.offset) / (self.factor + 1)
        if d >= self.aux_steps + 1:
            return d - self.aux_steps
        return d / (self.aux_steps + 1)

    def parse_sampling_schedule(self, spec: Union[None, str, Sequence[Number]], warn=lambda m: None) -> List[Number]:
        """`sampling_schedule` setter (:245-333)."""
        import numpy as np

        name = spec
        if spec is None:
            spec = list(range(self.num_timesteps))
        if isinstance(spec, str):
            base = [0] + list(self.dynamical_steps.keys())
            art = list(self.artificial_interpolation_steps.keys())
            if "only_dynamics" in spec:
                picked: list = []
                if "only_dynamics_plus" in spec:
                    n_extra = int(spec.replace("only_dynamics_plus", "").replace("_discrete", ""))
                    picked = list(np.linspace(0, base[1], n_extra + 1, endpoint=False))
                    if "_discrete" in spec:
                        picked = [int(np.floor(v)) for v in picked]
                else:
                    assert spec == "only_dynamics", f"Invalid sampling schedule: {spec}"
            elif spec.startswith("every"):
                nth = int(spec.replace("every", "").replace("th", "").replace("nd", "").replace("rd", ""))
                assert 1 <= nth <= self.num_timesteps, f"Invalid sampling schedule: {spec}"
                picked = art[::nth]
            elif spec.startswith("first"):
                first = float(spec.replace("first", "").replace("v2", ""))
                if first < 1:
                    assert 0 < first < 1, f"Invalid sampling schedule: {spec}, must end with number/float > 0"
                    picked = art[: int(np.ceil(first * len(art)))]
                else:
                    assert first.is_integer(), f"If first_n >= 1, it must be an integer, but got {first}"
                    assert 1 <= first <= self.num_timesteps, f"Invalid sampling schedule: {spec}"
                    picked = art[: int(first)]
            else:
                raise ValueError(f"Invalid sampling schedule: ``{spec}``. ")
            spec = sorted(set(picked + base))
        sched = list(spec)
        assert 1 <= sched[-1] <= self.num_timesteps, \
            f"Invalid sampling schedule: {sched}, must end with number/float <= {self.num_timesteps}"
        if sched[0] != 0:
            warn(f"Sampling schedule {name} must start at 0. Adding 0 to the beginning of it.")
            sched = [0] + sched
        if sched[-1] != self.num_timesteps - 1:
            warn(f"Are you sure you don't want to sample at the last timestep? (current last timestep: {sched[-1]})")
        for a, b in zip(sched, sched[1:]):
            assert b > a, f"Invalid sampling schedule not monotonically increasing: {sched}"
        if all(float(v).is_integer() for v in sched):
            sched = [int(v) for v in sched]
        return sched
