"""Result files for batched solutions.

Reference: ``save / load / save_result(sol, name, dir=None, f='xlsx')`` (Solverz/utilities/io.py:9-90).  The
reference writes one trajectory per workbook: a ``Time`` sheet plus one sheet per variable with columns
``<var>_<j>`` (io.py:29-42), the events into ``event_in_<name>.xlsx`` (44-60), or one MATLAB struct ``sol`` with
``Time / Y.<var> / te / ie / ye.<var>`` (62-79); an ``aesol`` becomes one sheet per variable (81-88).

A batch of scenarios does not fit a workbook, so this module adds ``f='npz'`` (the default here): the same names
as the MATLAB layout, flat keys ``Time``, ``Y.<var>``, ``te``, ``ie``, ``ye.<var>``, ``y.<var>`` and the solver
statistics under ``stats.<counter>``, every array keeping its leading scenario axis ``[B, ...]`` when the solution
is batched.  ``f='matlab'`` writes the reference's struct (with the scenario axis in front for a batch);
``f='xlsx'`` follows the reference sheet by sheet and, like the reference, handles one trajectory and needs
``openpyxl``.  Device tensors are brought to the host here; nothing in this file is on the hot path."""
from __future__ import annotations

import pickle
from pathlib import Path

import numpy as np

from .solvers.solution import aesol, daesol
from .variables import TimeVars, Vars

__all__ = ["save", "load", "save_result", "load_result"]

_STAT_KEYS = ("nstep", "nfeval", "nJeval", "ndecomp", "nreject", "nsolve", "nevent", "n_rows", "flag", "succeed")


def save(obj, filename: str):
    """io.py:9-11 (the reference pickles with dill; plain pickle here, lambdas are not supported)."""
    with open(filename, "wb") as file:
        pickle.dump(obj, file)


def load(filename: str):
    """io.py:14-16."""
    with open(filename, "rb") as file:
        return pickle.load(file)


def _host(a):
    if a is None:
        return None
    if hasattr(a, "detach"):                       # torch tensor left on the device by the batched solvers
        return a.detach().cpu().numpy()
    if isinstance(a, (Vars, TimeVars)):
        return a.array
    return np.asarray(a)


def _by_var(prefix: str, v, out: dict, slices=None):
    """Named slices of a ``Vars/TimeVars`` (last axis), else the flat array under ``prefix``."""
    if isinstance(v, (Vars, TimeVars)):
        for name in v.var_list:
            out[f"{prefix}.{name}"] = v[name]
    elif v is not None:
        arr = _host(v)
        if slices:
            for name, sl in slices.items():
                out[f"{prefix}.{name}"] = arr[..., sl]
        else:
            out[prefix] = arr


def _collect(sol, slices=None) -> dict:
    out = {}
    if isinstance(sol, daesol):
        out["Time"] = _host(sol.T)
        _by_var("Y", sol.Y, out, slices)
        if sol.ie is not None and np.size(_host(sol.ie)) > 0:
            out["te"], out["ie"] = _host(sol.te), _host(sol.ie)
            _by_var("ye", sol.ye, out, slices)
    elif isinstance(sol, aesol):
        _by_var("y", sol.y, out, slices)
    else:
        raise TypeError(f"Unknown solution type {type(sol)}")
    st = getattr(sol, "stats", None)
    for k in _STAT_KEYS:
        val = getattr(st, k, None)
        if val is not None:
            out[f"stats.{k}"] = _host(val)
    return out


def _nest(flat: dict) -> dict:
    d = {}
    for k, v in flat.items():
        head, _, tail = k.partition(".")
        if tail:
            d.setdefault(head, {})[tail] = v
        else:
            d[head] = v
    return d


def save_result(sol, name: str, dir: str = None, f: str = "npz", var_slices: dict = None):
    """Write ``sol`` to ``<dir>/<name>.{npz,mat,xlsx}`` and return the path.

    ``var_slices`` (name -> slice, e.g. ``mdl.var_slices``) names the columns of a solution that was computed from
    a plain array / device tensor instead of a ``Vars``."""
    base = Path(dir) if dir is not None else Path(".")
    flat = _collect(sol, var_slices)
    if f == "npz":
        path = base / f"{name}.npz"
        np.savez_compressed(path, **flat)
        return path
    if f == "matlab":
        from scipy.io import savemat
        path = base / f"{name}.mat"
        savemat(str(path), {"sol": _nest(flat)})
        return path
    if f == "xlsx":
        import pandas as pd
        try:
            import openpyxl  # noqa: F401
        except ImportError as e:
            raise ImportError("save_result(f='xlsx') needs openpyxl, like the reference; use f='npz' or "
                              "f='matlab'") from e
        lead = flat.get("Time", next(iter(flat.values())))
        if isinstance(sol, daesol) and np.ndim(lead) != 1:
            raise ValueError("a workbook holds one trajectory: save a batch with f='npz' or f='matlab'")
        path = base / f"{name}.xlsx"
        with pd.ExcelWriter(path, engine="openpyxl") as writer:
            if isinstance(sol, daesol):
                pd.DataFrame({"T": flat["Time"]}).to_excel(writer, sheet_name="Time")
                for k, arr in flat.items():
                    if k.startswith("Y."):
                        var = k[2:]
                        pd.DataFrame({f"{var}_{j}": arr[:, j] for j in range(arr.shape[1])}).to_excel(
                            writer, sheet_name=var)
            else:
                for k, arr in flat.items():
                    if k.startswith("y."):
                        if np.ndim(arr) != 1:
                            raise ValueError("a workbook holds one solution: save a batch with f='npz'")
                        pd.DataFrame({k[2:]: arr}).to_excel(writer, sheet_name=k[2:])
        if "ie" in flat:
            with pd.ExcelWriter(base / f"event_in_{name}.xlsx", engine="openpyxl") as writer:
                pd.DataFrame({"te": flat["te"]}).to_excel(writer, sheet_name="Event Time")
                pd.DataFrame({"ie": flat["ie"]}).to_excel(writer, sheet_name="Event Index")
                for k, arr in flat.items():
                    if k.startswith("ye."):
                        var = k[3:]
                        pd.DataFrame({f"{var}_{j}": arr[:, j] for j in range(arr.shape[1])}).to_excel(
                            writer, sheet_name=var)
        return path
    raise ValueError(f"unknown format {f!r}: 'npz', 'matlab' or 'xlsx'")


def load_result(path) -> dict:
    """Read a file written by ``save_result(f='npz')`` back into the nested layout of the MATLAB struct."""
    with np.load(path) as z:
        return _nest({k: z[k] for k in z.files})
