"""Link: one MCMC sample (tinyDA/link.py:1-48) and lazy link sequences.

The engine keeps histories structure-of-arrays on the device; ``LinkSequence`` is the host
view that ``sample()`` puts into the result dict where the reference puts a Python list of
Link objects.  It materialises ``Link`` objects on ``__getitem__`` / iteration, supports
slicing (``chain[burnin:]``) and ``+`` concatenation (both used in the reference's example
notebooks), and exposes the underlying arrays (``.parameters``, ``.prior`` ...) so that
``get_samples`` / ``to_inference_data`` never have to build one Python object per sample.
"""
import numpy as np


class Link:
    """Same attributes as the reference's Link (tinyDA/link.py:38-48)."""

    __slots__ = ("parameters", "prior", "model_output", "likelihood", "qoi", "posterior")

    def __init__(self, parameters, prior, model_output, likelihood, qoi=None):
        self.parameters = parameters
        self.prior = prior
        self.model_output = model_output
        self.likelihood = likelihood
        self.qoi = qoi
        self.posterior = self.prior + self.likelihood


class LinkSequence:
    """Array-backed, list-like sequence of Links for ONE chain at ONE level.

    parameters [n, d]; prior [n]; likelihood [n]; model_output [n, m] or None (not stored);
    accepted [n] bool or None.
    """

    def __init__(self, parameters, prior, likelihood, model_output=None, accepted=None):
        self.parameters = parameters
        self.prior = prior
        self.likelihood = likelihood
        self.model_output = model_output
        self.accepted = accepted

    def __len__(self):
        return self.parameters.shape[0]

    def _link(self, i):
        mo = None if self.model_output is None else self.model_output[i]
        return Link(self.parameters[i], float(self.prior[i]), mo, float(self.likelihood[i]), None)

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return LinkSequence(
                self.parameters[idx], self.prior[idx], self.likelihood[idx],
                None if self.model_output is None else self.model_output[idx],
                None if self.accepted is None else self.accepted[idx])
        n = len(self)
        if idx < 0:
            idx += n
        if not 0 <= idx < n:
            raise IndexError("link index out of range")
        return self._link(idx)

    def __iter__(self):
        for i in range(len(self)):
            yield self._link(i)

    def __add__(self, other):
        if not isinstance(other, LinkSequence):
            return list(self) + list(other)
        cat = np.concatenate
        mo = None
        if self.model_output is not None and other.model_output is not None:
            mo = cat([self.model_output, other.model_output])
        acc = None
        if self.accepted is not None and other.accepted is not None:
            acc = cat([self.accepted, other.accepted])
        return LinkSequence(cat([self.parameters, other.parameters]), cat([self.prior, other.prior]),
                            cat([self.likelihood, other.likelihood]), mo, acc)

    @property
    def posterior(self):
        return self.prior + self.likelihood
