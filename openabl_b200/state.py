"""Agent record layouts shared by tests, bench and the oracle fixtures.

A model's agent types are read from its `.abl` source; the host AoS layout is the natural
C struct layout the generated code (and the reference `c` backend) uses.
"""
import re

import numpy as np

_AGENT_RE = re.compile(r"agent\s+(\w+)\s*\{([^}]*)\}", re.S)
_MEMBER_RE = re.compile(r"(position\s+)?(\w+)\s+(\w+)\s*;")


def _strip_comments(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def parse_agents(abl_text):
    """-> list of (agent name, [(member name, type name, is_position)])"""
    out = []
    for m in _AGENT_RE.finditer(_strip_comments(abl_text)):
        members = [(mm.group(3), mm.group(2), bool(mm.group(1))) for mm in _MEMBER_RE.finditer(m.group(2))]
        out.append((m.group(1), members))
    return out


def agent_dtype(members, use_float=False):
    """numpy structured dtype with C layout for one agent type."""
    real = np.float32 if use_float else np.float64
    fields = []
    for name, ty, _ in members:
        if ty == "bool":
            fields.append((name, np.bool_))
        elif ty == "int":
            fields.append((name, np.int32))
        elif ty == "float":
            fields.append((name, real))
        elif ty == "float2":
            fields.append((name, real, (2,)))
        elif ty == "float3":
            fields.append((name, real, (3,)))
        else:
            raise ValueError("unsupported member type %s" % ty)
    return np.dtype(fields, align=True)


def read_raw(path, dtypes):
    """Reads a `<save path>.bin` dump: per agent type `u64 n, u32 stride`, then records.
    `dtypes` is the list of structured dtypes in declaration order."""
    out = []
    with open(path, "rb") as f:
        for dt in dtypes:
            n = int(np.frombuffer(f.read(8), dtype=np.uint64)[0])
            stride = int(np.frombuffer(f.read(4), dtype=np.uint32)[0])
            if stride != dt.itemsize:
                raise ValueError("record size mismatch: file %d, dtype %d" % (stride, dt.itemsize))
            out.append(np.frombuffer(f.read(n * stride), dtype=dt).copy())
    return out


# Absolute floor of single-precision comparisons, in ulps of the member's largest magnitude: a
# component of a unit vector that passes through zero keeps the absolute rounding error of unit-size
# operands (observed against the reference: up to ~100 float ulps of the scale after 10 steps, the
# reference evaluating literals in double and the kernels in float), so no relative bound can hold
# there.  The bar for use_float is therefore |a-b| <= max(1e-4 |b|, 128 ulp_f32(scale)); double
# precision comparisons use no floor at all.
F32_FLOOR_ULPS = 128


def max_rel_error(a, b, abs_floor=0.0, floor_ulps=0):
    """max |a-b| / |b| over all floating members of two structured arrays — a RELATIVE error
    (north_star: 1e-9 relative in double, 1e-4 with use_float), also for values far below 1 such
    as boids velocities.  Differences up to an absolute floor count as agreement: `abs_floor`
    (default 0: none do) or, per member, `floor_ulps` units in the last place of the member's
    largest magnitude in the member's own precision — single-precision callers pass a few ulps,
    because a float coordinate that happens to lie near zero in a world 44 units wide carries the
    rounding error of its operands' scale, not of its own; double-precision callers pass nothing.
    A remaining difference against an exact zero is reported as infinite."""
    worst = 0.0
    for name in a.dtype.names:
        x, y = a[name], b[name]
        if not np.issubdtype(x.dtype, np.floating):
            continue
        eps = float(np.finfo(x.dtype).eps)
        x = x.astype(np.float64)
        y = y.astype(np.float64)
        if not x.size:
            continue
        floor = max(abs_floor, floor_ulps * eps * float(np.abs(y).max()))
        d = np.abs(x - y)
        d[d <= floor] = 0.0
        nz = d > 0
        if nz.any():
            with np.errstate(divide="ignore"):
                worst = max(worst, float((d[nz] / np.abs(y[nz])).max()))
    return worst


def exact_members_equal(a, b):
    """True if all integer / bool members agree exactly."""
    for name in a.dtype.names:
        if not np.issubdtype(a[name].dtype, np.floating):
            if not np.array_equal(a[name], b[name]):
                return False
    return True
