"""Build oracle/_ref/: the UNMODIFIED reference hot path, byte-compiled from the sources where they lie.

TEST / BASELINE INFRASTRUCTURE ONLY (never imported by the product package).

The reference (GenomeNet/Genome-NAS) is pure Python: there is nothing to link.  Its "compiled form" is CPython
bytecode, so this recipe runs `py_compile` on the reference modules of the hot path directly from
/root/reference (read-only; no source is copied into the repository) and writes sourceless bytecode modules
(`*.gnasref`: the gpurun snapshot drops `*.pyc`) into oracle/_ref/ (git-ignored, NOT gpurun-ignored: it travels
to the GPU box like the built libgnas.so).  `load()` installs a meta-path finder that imports them.
`bench.py --impl reference` and the `cpu_baseline` / `gpu_eager_baseline` legs import the reference from
there and run ITS OWN train() / Architect / model classes; when oracle/_ref/ is absent they fall back to the
restatement in oracle/genomenas_oracle.py (`kind: "port"`).

    python oracle/build_ref.py            # needs /root/reference (authoring container only)

Modules = the transitive import closure of `model_search`, `darts_tools.architect`,
`generalNAS_tools.train_and_validate` and `darts_tools.discard_operations` inside the reference tree
(SURVEY.md section 8a/8c), plus the derived-architecture / mask-supernet modules used to make goldens.
"""
import importlib.util
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REFERENCE = os.environ.get("GNAS_REFERENCE", "/root/reference")

MODULES = [
    "model_search.py",
    "darts_tools/__init__.py",
    "darts_tools/architect.py",
    "darts_tools/auxiliary_functions.py",
    "darts_tools/cnn_eval.py",
    "darts_tools/comp_aux.py",
    "darts_tools/discard_operations.py",
    "darts_tools/genotype_parser.py",
    "darts_tools/model_discCNN.py",
    "generalNAS_tools/__init__.py",
    "generalNAS_tools/custom_loss_functions.py",
    "generalNAS_tools/genotypes.py",
    "generalNAS_tools/operations_14_9.py",
    "generalNAS_tools/train_and_validate.py",
    "generalNAS_tools/utils.py",
]


EXT = ".gnasref"


def available():
    return os.path.isfile(os.path.join(OUT, "model_search" + EXT))


def build(force=False):
    """Returns True when oracle/_ref/ is usable afterwards."""
    if not os.path.isdir(REFERENCE):
        return available()          # GPU box / fresh clone: use whatever travelled with the snapshot
    stamp = os.path.join(OUT, ".stamp")
    want = sys.version + "".join(f"{m}:{os.path.getmtime(os.path.join(REFERENCE, m))};" for m in MODULES)
    if not force and available() and os.path.exists(stamp) and open(stamp).read() == want:
        return True
    for rel in MODULES:
        src = os.path.join(REFERENCE, rel)
        dst = os.path.join(OUT, rel[:-3] + EXT)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(src, cfile=dst, dfile=rel, doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(stamp, "w") as fh:
        fh.write(want)
    return True


class _RefFinder:
    """meta-path finder: module `a.b` <- oracle/_ref/a/b.gnasref (or a/b/__init__.gnasref), sourceless bytecode"""

    @staticmethod
    def find_spec(name, path=None, target=None):
        from importlib.machinery import SourcelessFileLoader
        from importlib.util import spec_from_file_location
        base = os.path.join(OUT, *name.split("."))
        if os.path.isfile(base + EXT):
            return spec_from_file_location(name, base + EXT, loader=SourcelessFileLoader(name, base + EXT))
        init = os.path.join(base, "__init__" + EXT)
        if os.path.isfile(init):
            return spec_from_file_location(name, init, loader=SourcelessFileLoader(name, init),
                                           submodule_search_locations=[base])
        return None


def load():
    """Import the byte-compiled reference.  Returns a namespace with the reference modules, or None when
    oracle/_ref/ does not exist.  The only shim is an empty `matplotlib` module (generalNAS_tools/utils.py:38
    imports it; it is not installed in this image and never used on this path)."""
    if not available():
        return None
    import types
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    if not any(f is _RefFinder for f in sys.meta_path):
        sys.meta_path.insert(0, _RefFinder)
    import model_search
    from darts_tools import architect, discard_operations, model_discCNN
    from generalNAS_tools import train_and_validate
    ns = types.SimpleNamespace(model_search=model_search, architect=architect, model_discCNN=model_discCNN,
                               train_and_validate=train_and_validate, discard_operations=discard_operations)
    for mod in (model_search, architect, model_discCNN, train_and_validate):
        f = getattr(mod, "__file__", "") or ""
        if not f.startswith(OUT):
            raise RuntimeError(f"reference module {mod.__name__} was imported from {f}, not from oracle/_ref")
    return ns


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref", "ready" if ok else "unavailable (no reference checkout)")
