"""Cost weights of the reference's ``CostFunction("WX1")`` (planners/common/cost/cost_function.py:5-12).

Only the weights live on the host; ``cost_total`` (:41-50) is evaluated inside the CUDA kernel
(warp-shuffle reduction per candidate).  ``as_device_weights`` is what the planners marshal into
``fiss_params``.
"""


class CostFunction:
    def __init__(self, cost_type: str):
        if cost_type == "WX1":
            self.w_T = 10
            self.w_V = 1
            self.w_A = 0.1
            self.w_J = 0.1
            self.w_D = 0.1
            self.w_LC = 10
        else:
            raise ValueError(f"unknown cost type {cost_type!r}")
        # cost_total uses the literal 10.0 - t[-1], not w_T or max_t (cost_function.py:42)
        self.time_offset = 10.0

    def as_device_weights(self):
        return dict(time_offset=self.time_offset, w_V=float(self.w_V), w_A=float(self.w_A),
                    w_J=float(self.w_J), w_LC=float(self.w_LC))
