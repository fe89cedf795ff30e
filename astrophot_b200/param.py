"""Parameter DAG (host-side bookkeeping around the hot path).

Mirrors the public surface of the reference's ``astrophot/param`` package
(`param/base.py:10-197`, `param/parameter.py:18-760`, `param/param_context.py`)
— ``Parameter_Node`` leaves hold values, branches hold structure, a node whose
value is another node is a *pointer* (this is how joint multi-band fits share
``center``/``q``/… between models, `parameter.py:490-496`) — but it is an
independent implementation: values live on the host as float64 torch tensors,
and the device only ever sees the flat vector + slot table that
``lowering.py`` builds from this graph.
"""
from collections import OrderedDict
from types import FunctionType

import numpy as np
import torch

from . import AP_config
from .errors import InvalidParameter

__all__ = ["Node", "Parameter_Node", "Param_Unlock", "Param_SoftLimits", "Param_Mask"]

_HOST = "cpu"


def _t(x):
    return torch.as_tensor(x, dtype=torch.float64, device=_HOST)


# ---------------------------------------------------------------------------
# value <-> representation maps (reference: utils/conversions/optimization.py:6-54)
# ---------------------------------------------------------------------------
def boundaries(val, limits):
    """value inside ``limits`` -> unbounded representation."""
    val = _t(val)
    lo, hi = limits
    if lo is None:
        return val - 1.0 / (val - hi)
    if hi is None:
        return val - 1.0 / (val - lo)
    return torch.tan((val - lo) * np.pi / (hi - lo) - np.pi / 2)


def inv_boundaries(rep, limits):
    """unbounded representation -> value inside ``limits``."""
    rep = _t(rep)
    lo, hi = limits
    if lo is None:
        return (rep + hi - torch.sqrt((rep - hi) ** 2 + 4)) * 0.5
    if hi is None:
        return (rep + lo + torch.sqrt((rep - lo) ** 2 + 4)) * 0.5
    return (torch.arctan(rep) + np.pi / 2) * (hi - lo) / np.pi + lo


def cyclic_boundaries(val, limits):
    val = _t(val)
    lo, hi = limits
    return lo + torch.remainder(val - lo, hi - lo)


class Node:
    """A vertex of the parameter graph: a name, ordered children, a lock."""

    global_unlock = False

    def __init__(self, name, **kwargs):
        if ":" in name:
            raise ValueError(f"Node names must not have ':' character. Cannot use name: {name}")
        self.name = name
        self.nodes = OrderedDict()
        if "state" in kwargs:
            self.set_state(kwargs["state"])
            return
        if "link" in kwargs:
            self.link(*kwargs["link"])
        self.locked = kwargs.get("locked", False)

    # -- structure -----------------------------------------------------
    def link(self, *nodes):
        for node in nodes:
            below = node.flat(include_locked=True, include_links=True)
            if self.identity in below or node is self:
                raise InvalidParameter(
                    "Parameter structure must be Directed Acyclic Graph! Adding this node would create a cycle"
                )
            self.nodes[node.name] = node

    def unlink(self, *nodes):
        for node in nodes:
            del self.nodes[node.name]

    def dump(self):
        self.unlink(*list(self.nodes.values()))

    # -- state (plain dict; reference: `param/base.py:134-161`) ------------------------------------------------------
    def get_state(self):
        state = {"name": self.name, "identity": self.identity}
        if self.locked:
            state["locked"] = True
        if len(self.nodes) > 0:
            state["nodes"] = [n.get_state() for n in self.nodes.values()]
        return state

    _loading = None      # identity -> node built so far by the set_state call in progress

    def set_state(self, state):
        """(A node that hangs under several parents appears once per parent in the state, with one identity: it is
        rebuilt once and linked to all of them, so shared parameters stay shared.)"""
        root = Node._loading is None
        if root:
            Node._loading = {}
        try:
            self.name = state["name"]
            self._identity = state["identity"]
            Node._loading[state["identity"]] = self
            for sub in state.get("nodes", ()):
                known = Node._loading.get(sub["identity"])
                self.link(known if known is not None else self.__class__(name=sub["name"], state=sub))
            self.locked = state.get("locked", False)
        finally:
            if root:
                Node._loading = None

    @property
    def leaf(self):
        return len(self.nodes) == 0

    @property
    def branch(self):
        return len(self.nodes) > 0

    @property
    def identity(self):
        return getattr(self, "_identity", id(self))

    def __getitem__(self, key):
        if key == self.name:
            return self
        if key in self.nodes:
            return self.nodes[key]
        if isinstance(key, str) and ":" in key:
            head, rest = key.split(":", 1)
            if head == self.name:
                return self[rest]
            return self.nodes[head][rest]
        if isinstance(key, int):
            for node in self.nodes.values():
                if node.identity == key:
                    return node
                try:
                    return node[key]
                except KeyError:
                    pass
        raise KeyError(f"Unrecognized key for '{self.name}': {key}")

    def __contains__(self, key):
        return key in self.nodes

    def __eq__(self, other):
        return self is other

    __hash__ = object.__hash__

    def __iter__(self):
        return (n for n in self.nodes.values() if not n.locked)

    def flat(self, include_locked=True, include_links=False):
        """Ordered {identity: node} of the leaves reachable from here
        (depth first, first visit wins — this fixes the column order of J)."""
        out = OrderedDict()
        unlocked = include_locked or Node.global_unlock
        if self.leaf and self.value is not None:
            if (not self.locked) or unlocked:
                out[self.identity] = self
        for node in self.nodes.values():
            if node.locked and not unlocked:
                continue
            if node.leaf and node.value is not None:
                out[node.identity] = node
            else:
                if include_links and ((not node.locked) or unlocked):
                    out[node.identity] = node
                for k, v in node.flat(include_locked).items():
                    out.setdefault(k, v)
        return out

    def __str__(self):
        return f"Node: {self.name}"

    __repr__ = __str__


class Parameter_Node(Node):
    """Leaf = tensor value (+limits, cyclic flag, uncertainty, ``prof``);
    branch = collection; pointer = value is another node; function = value
    computed from other nodes."""

    def __init__(self, name, **kwargs):
        super().__init__(name, **kwargs)
        if "state" in kwargs:          # rebuilt by set_state (called from Node.__init__)
            return
        hold = self.locked
        self.locked = False
        self._value = None
        self._shape = None
        self._uncertainty = None
        self.prof = kwargs.get("prof", None)
        self.limits = kwargs.get("limits", [None, None])
        self.cyclic = kwargs.get("cyclic", False)
        self.shape = kwargs.get("shape", None)
        self.value = kwargs.get("value", None)
        self.units = kwargs.get("units", "none")
        self.uncertainty = kwargs.get("uncertainty", None)
        self.locked = hold

    # -- value ---------------------------------------------------------
    @property
    def value(self):
        v = self._value
        if isinstance(v, Parameter_Node):
            return v.value
        if isinstance(v, FunctionType):
            return v(self)
        return v

    @value.setter
    def value(self, val):
        if self.locked and not Node.global_unlock:
            return
        if val is None:
            self._value = None
            self._shape = None
            return
        if isinstance(val, str):
            self._value = val
            return
        if isinstance(val, Parameter_Node):
            self._value = val
            self._shape = None
            self.dump()
            self.link(val)
            return
        if isinstance(val, FunctionType):
            self._value = val
            self._shape = None
            return
        if len(self.nodes) > 0:
            self.vector_set_values(val)
            self._shape = None
            return
        self._set_leaf(val, soft=getattr(self, "_soft", False))
        self.dump()

    def _set_leaf(self, val, soft=False):
        val = _t(val).clone()
        if self._shape is not None:
            val = val.reshape(self._shape)
        else:
            self._shape = tuple(val.shape)
        lo, hi = self.limits
        if self.cyclic:
            val = lo + torch.remainder(val - lo, hi - lo)
        elif soft:
            if lo is not None:
                val = torch.maximum(val, lo + 1e-3)
            if hi is not None:
                val = torch.minimum(val, hi - 1e-3)
        else:
            if lo is not None and not bool(torch.all(val > lo)):
                raise InvalidParameter(f"{self.name} has lower limit {lo.tolist()}")
            if hi is not None and not bool(torch.all(val < hi)):
                raise InvalidParameter(f"{self.name} has upper limit {hi.tolist()}")
        self._value = val

    @property
    def shape(self):
        if isinstance(self._value, Parameter_Node):
            return self._value.shape
        if isinstance(self._value, FunctionType):
            return tuple(self.value.shape)
        return self._shape

    @shape.setter
    def shape(self, shape):
        self._shape = None if shape is None else tuple(shape)

    @property
    def prof(self):
        return self._prof

    @prof.setter
    def prof(self, prof):
        if self.locked and not Node.global_unlock:
            return
        self._prof = None if prof is None else _t(prof)

    @property
    def uncertainty(self):
        return self._uncertainty

    @uncertainty.setter
    def uncertainty(self, unc):
        if self.locked and not Node.global_unlock:
            return
        if unc is None:
            self._uncertainty = None
            return
        unc = _t(unc)
        if self.value is not None and not isinstance(self.value, str) and unc.numel() == 1:
            unc = unc * torch.ones_like(self.value)
        self._uncertainty = unc

    @property
    def limits(self):
        return self._limits

    @limits.setter
    def limits(self, limits):
        if self.locked and not Node.global_unlock:
            return
        lo = None if limits[0] is None else _t(limits[0])
        hi = None if limits[1] is None else _t(limits[1])
        self._limits = (lo, hi)

    # -- masks / identities -------------------------------------------
    @property
    def mask(self):
        if not self.leaf:
            return self.vector_mask()
        m = getattr(self, "_mask", None)
        if m is None:
            return torch.ones(self.shape, dtype=torch.bool)
        return m

    @property
    def size(self):
        if self.leaf:
            return int(self.value.numel())
        return int(sum(n.size for n in self.flat(False, False).values()))

    def __len__(self):
        return self.size

    @property
    def identities(self):
        if self.leaf:
            idstr = str(self.identity)
            return np.array([f"{idstr}:{i}" for i in range(self.size)])
        vec = [n.identities for n in self.flat(False, False).values()]
        return np.concatenate(vec) if vec else np.array(())

    @property
    def names(self):
        if self.leaf:
            if self.size == 1:
                return np.array([self.name])
            return np.array([f"{self.name}:{i}" for i in range(self.size)])
        vec = [n.names for n in self.flat(False, False).values()]
        return np.concatenate(vec) if vec else np.array(())

    # -- flat vector views --------------------------------------------
    def _leaves(self):
        return list(self.flat(include_locked=False, include_links=False).values())

    def _cat(self, pieces):
        pieces = list(pieces)
        if pieces:
            return torch.cat(pieces)
        return torch.zeros(0, dtype=torch.float64)

    def vector_values(self):
        if self.leaf:
            return self.value[self.mask].flatten()
        return self._cat(n.vector_values() for n in self._leaves())

    def vector_uncertainty(self):
        if self.leaf:
            if self._uncertainty is None:
                self._uncertainty = torch.ones_like(self.value)
            return self._uncertainty[self.mask].flatten()
        return self._cat(n.vector_uncertainty() for n in self._leaves())

    def vector_mask(self):
        if self.leaf:
            return self.mask.flatten()
        pieces = [n.vector_mask() for n in self._leaves()]
        return torch.cat(pieces) if pieces else torch.zeros(0, dtype=torch.bool)

    def vector_identities(self):
        if self.leaf:
            return self.identities[self.vector_mask().numpy()].flatten()
        vec = [n.vector_identities() for n in self._leaves()]
        return np.concatenate(vec) if vec else np.array(())

    def vector_names(self):
        if self.leaf:
            return self.names[self.vector_mask().numpy()].flatten()
        vec = [n.vector_names() for n in self._leaves()]
        return np.concatenate(vec) if vec else np.array(())

    def vector_representation(self):
        return self.vector_transform_val_to_rep(self.vector_values())

    def _split(self, vec):
        """Cut a masked flat vector into per-leaf pieces (O(leaves), unlike
        the reference's O(leaves^2) running mask sums, parameter.py:272-280)."""
        at = 0
        for node in self._leaves():
            n = int(node.mask.sum())
            yield node, vec[at : at + n]
            at += n

    def vector_set_values(self, values):
        values = _t(values).flatten()
        if self.leaf:
            self._value[self.mask] = values
            return
        for node, piece in self._split(values):
            node.vector_set_values(piece)

    def vector_set_uncertainty(self, uncertainty):
        uncertainty = _t(uncertainty).flatten()
        if self.leaf:
            if self._uncertainty is None:
                self._uncertainty = torch.ones_like(self.value)
            self._uncertainty[self.mask] = uncertainty
            return
        for node, piece in self._split(uncertainty):
            node.vector_set_uncertainty(piece)

    def vector_set_mask(self, mask):
        mask = torch.as_tensor(mask, dtype=torch.bool)
        if self.leaf:
            self._mask = mask.reshape(self.shape)
            return
        at = 0
        for node in self._leaves():
            node.vector_set_mask(mask[at : at + node.size])
            at += node.size

    def vector_set_representation(self, rep):
        self.vector_set_values(self.vector_transform_rep_to_val(rep))

    def vector_transform_rep_to_val(self, rep):
        rep = _t(rep)
        if self.leaf:
            if self.cyclic:
                return cyclic_boundaries(rep, self.limits)
            if self.limits[0] is None and self.limits[1] is None:
                return rep
            return inv_boundaries(rep, self._masked_limits())
        return self._cat(n.vector_transform_rep_to_val(p) for n, p in self._split(rep))

    def vector_transform_val_to_rep(self, val):
        val = _t(val)
        if self.leaf:
            if self.cyclic:
                return cyclic_boundaries(val, self.limits)
            if self.limits[0] is None and self.limits[1] is None:
                return val
            return boundaries(val, self._masked_limits())
        return self._cat(n.vector_transform_val_to_rep(p) for n, p in self._split(val))

    def _masked_limits(self):
        out = []
        for lim in self.limits:
            if lim is not None and lim.numel() > 1:
                lim = lim.reshape(self.shape)[self.mask].flatten()
            out.append(lim)
        return tuple(out)

    def to(self, dtype=None, device=None):
        return self

    # -- persistence / display ----------------------------------------
    def set_state(self, state):
        """Rebuild this node from ``get_state()`` of another (reference: `parameter.py:632-647`)."""
        for attr, default in (("_value", None), ("_shape", None), ("_uncertainty", None), ("_prof", None),
                              ("_limits", (None, None)), ("cyclic", False), ("units", "none")):
            if not hasattr(self, attr):
                object.__setattr__(self, attr, default)
        self.locked = False
        Node.set_state(self, state)
        hold = self.locked
        self.locked = False
        self.units = state.get("units", None)
        self.limits = state.get("limits", (None, None))
        self.cyclic = state.get("cyclic", False)
        if "shape" in state and "value" in state and not isinstance(state["value"], str):
            self.shape = state["shape"]
        if isinstance(state.get("value"), str):            # pointer / function place-holder: see relink()
            self._value = state["value"]
        else:
            self.value = state.get("value", None)
        self.uncertainty = state.get("uncertainty", None)
        self.prof = state.get("prof", None)
        self.locked = hold

    def relink(self):
        """After ``set_state`` on the root of a graph: turn the ``NODE:<identity>`` place-holders of pointer nodes back
        into pointers (identities travel with the state).  The reference saves these place-holders but never resolves
        them; here a saved joint fit keeps its shared parameters."""
        seen = {}

        def walk(n):
            if n.identity in seen:
                return
            seen[n.identity] = n
            for c in n.nodes.values():
                walk(c)

        walk(self)
        by_id = {str(k): v for k, v in seen.items()}
        for n in seen.values():
            v = getattr(n, "_value", None)
            if isinstance(v, str) and v.startswith("NODE:") and v[5:] in by_id:
                hold, n.locked = n.locked, False
                n._value = None
                n.value = by_id[v[5:]]
                n.locked = hold
        return self

    def get_state(self):
        state = Node.get_state(self)
        if isinstance(self._value, Parameter_Node):
            state["value"] = "NODE:" + str(self._value.identity)
        elif isinstance(self._value, FunctionType):
            state["value"] = "FUNCTION:" + self._value.__name__
        if self.leaf and self.value is not None:
            state["value"] = self.value.tolist()
            state["shape"] = list(self.shape)
            if self.units is not None:
                state["units"] = self.units
            if self._uncertainty is not None:
                state["uncertainty"] = self._uncertainty.tolist()
            if not (self.limits[0] is None and self.limits[1] is None):
                state["limits"] = [None if l is None else l.tolist() for l in self.limits]
            if self.cyclic:
                state["cyclic"] = True
            if self.prof is not None:
                state["prof"] = self.prof.tolist()
        return state

    # -- report (the text users read after a fit; same layout as the reference, `parameter.py:677-742`) ------------
    def print_params(self, include_locked=True, include_prof=True, include_id=True):
        """One line for this node: ``name: value +- uncertainty [units], limits: (lo, hi), cyclic, locked, prof: ...``
        for a leaf; ``name points to: <line of the target>`` for a pointer; ``name:`` for a branch or function node."""
        tag = f" (id-{self.identity})" if include_id else ""
        if isinstance(self._value, Parameter_Node):
            return f"{self.name}{tag} points to: " + self._value.print_params(include_locked, include_prof, include_id)
        if not self.leaf:
            if include_id:
                kind = f"function node, {self._value.__name__}" if isinstance(self._value, FunctionType) else "branch node"
                tag = f" (id-{self.identity}, {kind})"
            return f"{self.name}{tag}:\n"

        def listed(t):
            return None if t is None else t.detach().cpu().tolist()

        lo, hi = self.limits
        parts = [f"{self.name}{tag}: {listed(self.value)}"]
        if self.uncertainty is not None:
            parts.append(f" +- {listed(self.uncertainty)}")
        parts.append(f" [{self.units}]")
        if lo is not None or hi is not None:
            parts.append(f", limits: ({listed(lo)}, {listed(hi)})")
        if self.cyclic:
            parts.append(", cyclic")
        if self.locked:
            parts.append(", locked")
        if include_prof and self.prof is not None:
            parts.append(f", prof: {listed(self.prof)}")
        return "".join(parts)

    def __str__(self):
        head = self.print_params(include_locked=True, include_prof=False, include_id=False)
        if self.leaf or isinstance(self._value, Parameter_Node):
            return head
        return head + "\n".join(n.print_params(include_locked=True, include_prof=False, include_id=False)
                                for n in self.flat(include_locked=True, include_links=False).values())

    def __repr__(self, level=0, indent="  "):
        head = indent * level + self.print_params(include_locked=True, include_prof=False, include_id=True)
        if self.leaf or isinstance(self._value, Parameter_Node):
            return head
        return head + "\n".join(n.__repr__(level=level + 1, indent=indent) for n in self.nodes.values())


class Param_Unlock:
    """Temporarily lift the lock of one node (or of all nodes)."""

    def __init__(self, param=None):
        self.param = param

    def __enter__(self):
        if self.param is None:
            self.saved = Node.global_unlock
            Node.global_unlock = True
        else:
            self.saved = self.param.locked
            self.param.locked = False

    def __exit__(self, *a):
        if self.param is None:
            Node.global_unlock = self.saved
        else:
            self.param.locked = self.saved


class Param_SoftLimits:
    """Inside this context out-of-range values are clipped, not rejected."""

    def __init__(self, param):
        self.param = param

    def __enter__(self):
        self.param._soft = True

    def __exit__(self, *a):
        self.param._soft = False


class Param_Mask:
    """Temporarily hide elements of the flat vector (reference:
    `param_context.py:62-102`; used by chunked fits)."""

    def __init__(self, param, new_mask):
        self.param = param
        self.new_mask = torch.as_tensor(new_mask, dtype=torch.bool)

    def __enter__(self):
        self.old_mask = self.param.vector_mask()
        full = self.old_mask.clone()
        full[self.old_mask] = self.new_mask
        self.param.vector_set_mask(full)

    def __exit__(self, *a):
        self.param.vector_set_mask(self.old_mask)
