#!/bin/sh
# Private, git-ignored copy of the reference under baseline/_ref: the reference is a
# pure-Python package, so "installing" it is a copy plus the two compatibility edits below.  Runs only where
# /root/reference exists (the build container); the copy travels to the GPU box with the snapshot.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
[ -d "$REF/Solverz" ] || { echo "no reference at $REF"; exit 0; }
rm -rf "$HERE/_ref"
mkdir -p "$HERE/_ref"
cp -r "$REF/Solverz" "$HERE/_ref/"
# sympy 1.14: Derivative(Function('F'), x) raises (the reference pins sympy<=1.13.3); the LHS is display-only
sed -i "s/Derivative(Function('F'), diff_var)/Derivative(Function('F')(diff_var), diff_var)/" "$HERE/_ref/Solverz/equation/eqn.py"
# setuptools_scm artefact that a raw checkout lacks
echo "__version__ = '0.8.5'" > "$HERE/_ref/Solverz/_version.py"
find "$HERE/_ref" -name "__pycache__" -type d -exec rm -rf {} + 2>/dev/null || true
echo "baseline/_ref ready"
