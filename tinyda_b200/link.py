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

    def __init__(self, parameters, prior, likelihood, model_output=None, accepted=None, qoi=None):
        self.parameters = parameters
        self.prior = prior
        self.likelihood = likelihood
        self.model_output = model_output
        self.accepted = accepted
        self.qoi = qoi

    def __len__(self):
        return self.parameters.shape[0]

    def _link(self, i):
        mo = None if self.model_output is None else self.model_output[i]
        q = None if self.qoi is None else self.qoi[i]
        return Link(self.parameters[i], float(self.prior[i]), mo, float(self.likelihood[i]), q)

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return LinkSequence(
                self.parameters[idx], self.prior[idx], self.likelihood[idx],
                None if self.model_output is None else self.model_output[idx],
                None if self.accepted is None else self.accepted[idx],
                None if self.qoi is None else self.qoi[idx])
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
        q = None
        if self.qoi is not None and other.qoi is not None:
            q = cat([self.qoi, other.qoi])
        return LinkSequence(cat([self.parameters, other.parameters]), cat([self.prior, other.prior]),
                            cat([self.likelihood, other.likelihood]), mo, acc, q)

    @property
    def posterior(self):
        return self.prior + self.likelihood


class CompactHistory:
    """Finest-level Link history of all local chains as the engine ships it: per block of records the
    accept flag of every record plus the fields of the ACCEPTED records only.  A rejected step
    re-appends the same Link object in the reference (chain.py:116, :434; proposal.py:1601), so record
    r of a chain is its latest accepted record at or before r; ``chain(c)`` expands one chain to a
    LinkSequence, ``dense(field)`` all of them at once."""

    FIELDS = ("theta", "prior", "like", "output", "qoi")

    def __init__(self, n_chains):
        self.n_chains = int(n_chains)
        self.chunks = []

    def append(self, chunk):
        self.chunks.append(chunk)

    @property
    def n_records(self):
        return sum(ch.nrec for ch in self.chunks)

    def n_rows(self):
        return sum(int(ch.offsets[-1]) for ch in self.chunks)

    def _rows(self, c, field):
        parts = []
        for ch in self.chunks:
            a = getattr(ch, field)
            if a is None:
                return None
            parts.append(a[int(ch.offsets[c]):int(ch.offsets[c + 1])])
        return parts[0] if len(parts) == 1 else np.concatenate(parts)

    def accepted(self, c):
        acc = [ch.accept[:, c] for ch in self.chunks]
        return acc[0] if len(acc) == 1 else np.concatenate(acc)

    def chain(self, c):
        flags = self.accepted(c)
        idx = np.cumsum(flags, dtype=np.int64) - 1           # record -> row of the chain (first record is a row)
        take = lambda rows: None if rows is None else rows[idx]
        acc = flags.astype(bool)
        acc[0] = False                                       # the initial Link is not an accepted proposal
        return LinkSequence(take(self._rows(c, "theta")), take(self._rows(c, "prior")), take(self._rows(c, "like")),
                            take(self._rows(c, "output")), acc, take(self._rows(c, "qoi")))

    def dense(self, field, burnin=0):
        """All chains at once: [n_chains, n_records - burnin, width] (width dropped for prior / like)."""
        C = self.n_chains
        outs = []
        carry = None                                          # last row of every chain from the previous chunks
        for ch in self.chunks:
            rows = getattr(ch, field)
            if rows is None:
                raise ValueError("field %r was not stored" % field)
            cnt = np.cumsum(ch.accept, axis=0, dtype=np.int64)        # [nrec, C] accepted so far in this chunk
            pos = ch.offsets[:-1][None, :] + cnt - 1                  # row of (record, chain); offsets-1 where cnt == 0
            x = rows[np.maximum(pos, 0)]
            if carry is not None:
                none_yet = cnt == 0
                if none_yet.any():
                    r, c = np.nonzero(none_yet)
                    x[r, c] = carry[c]
            last = ch.offsets[1:] - 1
            has = ch.offsets[1:] > ch.offsets[:-1]
            new_carry = rows[np.maximum(last, 0)]
            carry = new_carry if carry is None else np.where(has.reshape((-1,) + (1,) * (rows.ndim - 1)), new_carry, carry)
            outs.append(x)
        x = outs[0] if len(outs) == 1 else np.concatenate(outs, axis=0)
        x = x[burnin:]
        return np.swapaxes(x, 0, 1)


class SampleResult(dict):
    """The dict ``sample()`` returns (sampler.py:305-309, :406-439, :510-547).  The per-chain entries
    (``chain_i`` / ``chain_fine_i`` / ``chain_coarse_i`` / ``chain_l{l}_i``) are virtual until they are
    read: a run with tens of thousands of chains does not build one Python object per chain up front.
    Reading a key, ``in``, ``len``, ``keys`` / ``items`` / ``values`` and iteration behave like the plain
    dict of the reference; ``history`` holds the finest level of the local chains in compacted form
    (``CompactHistory``) for bulk post-processing."""

    def __init__(self, info=()):
        super().__init__(info)
        self._families = []          # (prefix, lo, hi, factory(local index) -> value)
        self.history = None
        self.local_chains = None

    def add_chains(self, fmt, lo, hi, factory):
        self._families.append((fmt.format(""), int(lo), int(hi), factory))

    def _family(self, key):
        if isinstance(key, str):
            for prefix, lo, hi, factory in self._families:
                tail = key[len(prefix):]
                if key.startswith(prefix) and tail.isdigit() and lo <= int(tail) < hi:
                    return int(tail) - lo, factory
        return None

    def __missing__(self, key):
        hit = self._family(key)
        if hit is None:
            raise KeyError(key)
        v = hit[1](hit[0])
        dict.__setitem__(self, key, v)
        return v

    def __contains__(self, key):
        return dict.__contains__(self, key) or self._family(key) is not None

    def keys(self):
        ks = [k for k in dict.keys(self) if self._family(k) is None]
        for prefix, lo, hi, _ in self._families:
            ks.extend(prefix + str(i) for i in range(lo, hi))
        return ks

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self.keys())

    def get(self, key, default=None):
        return self[key] if key in self else default

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]
