"""dgl.function builtin reducers (test stand-in, see dgl/__init__.py)."""
import torch


class BuiltinReducer:
    def __init__(self, name, msg_field, out_field):
        self.name, self.msg_field, self.out_field = name, msg_field, out_field

    def reduce(self, m, dst, n):
        width = 1
        for w in m.shape[1:]:
            width *= int(w)
        flat = m.reshape(m.shape[0], width)
        if self.name in ("sum", "mean"):
            # dense incidence product (NOT index_add_: the oracle uses that)
            inc = torch.zeros(n, m.shape[0], dtype=m.dtype, device=m.device)
            if m.shape[0]:
                inc[dst, torch.arange(m.shape[0], device=m.device)] = 1
            out = inc @ flat
            if self.name == "mean":
                deg = inc.sum(1).clamp(min=1)
                out = out / deg[:, None]
        else:
            # per-node loop over the in-edges; nodes without in-edges keep zeros
            rows = []
            for u in range(n):
                sel = flat[dst == u]
                if sel.shape[0] == 0:
                    rows.append(flat.new_zeros(flat.shape[1]))
                else:
                    rows.append(sel.max(0).values if self.name == "max" else sel.min(0).values)
            out = torch.stack(rows) if rows else flat.new_zeros((0, flat.shape[1]))
        return out.reshape((n,) + tuple(m.shape[1:]))


def sum(msg, out):  # noqa: A001
    return BuiltinReducer("sum", msg, out)


def mean(msg, out):
    return BuiltinReducer("mean", msg, out)


def max(msg, out):  # noqa: A001
    return BuiltinReducer("max", msg, out)


def min(msg, out):  # noqa: A001
    return BuiltinReducer("min", msg, out)
