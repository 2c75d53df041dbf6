"""solid::filter -- the `Filter<I, O>` plugin trait (filter/mod.rs:9-22) and its implementors."""
from __future__ import annotations


class Filter:
    """trait Filter<I, O> (filter/mod.rs:9-22): execute / execute_block / frequency_response /
    group_delay.  `execute*` run on the GPU through the C ABI; the two analysis methods are
    host-side f64 arithmetic over the stored coefficients, as in the reference."""

    def execute(self, sample):
        """Filter::execute -- one input sample (per channel) in, Vec<O> out."""
        return self.execute_block([sample] if self.n_channels == 1 else
                                  [[s] for s in sample])

    def execute_block(self, samples):
        raise NotImplementedError

    def frequency_response(self, frequency: float) -> complex:
        raise NotImplementedError

    def group_delay(self, frequency: float) -> float:
        raise NotImplementedError


from . import group_delay  # noqa: E402,F401
