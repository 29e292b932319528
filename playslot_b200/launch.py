"""
Run an unmodified reference script with its `models` package replaced by the playslot_b200 mirror:

    cd /path/to/PlaySlot && python -m playslot_b200.launch src/05_evaluate_PlaySlot.py -d experiments/... [args]

The reference scripts are started as `python src/<script>.py`, i.e. with `src/` first on sys.path, and do
`from models.SAVi import SAVi` etc. (lib/setup_model.py:49-111).  This launcher puts `src/` on sys.path the same
way, registers the mirror under the top-level name `models` before the script starts, and appends the
reference's own `src/models` directories to the mirror's package search paths so that sub-modules the mirror
does not provide (e.g. `models.Downstream`) still resolve to the reference files.
"""
import importlib
import os
import runpy
import sys

_SUBPACKAGES = ("BlocksUtils", "EncodersDecoders", "Predictors")
_MODULES = ("SAVi",
            "BlocksUtils.attention", "BlocksUtils.initializers", "BlocksUtils.model_blocks", "BlocksUtils.model_utils",
            "BlocksUtils.quantization", "BlocksUtils.transition_models",
            "EncodersDecoders.encoders", "EncodersDecoders.decoders",
            "Predictors.LatentAction", "Predictors.DynamicsModels", "Predictors.LatentActionPredictors",
            "Predictors.ActionConditional_OCVP", "Predictors.transformers", "Predictors.predictor_wrappers")


def install(reference_src=None):
    """ alias playslot_b200.models[.*] as models[.*]; returns the list of aliased module names """
    mirror = importlib.import_module("playslot_b200.models")
    sys.modules["models"] = mirror
    names = ["models"]
    for sp in _SUBPACKAGES:
        m = importlib.import_module(f"playslot_b200.models.{sp}")
        sys.modules[f"models.{sp}"] = m
        names.append(f"models.{sp}")
        if reference_src:
            ref_dir = os.path.join(reference_src, "models", sp)
            if os.path.isdir(ref_dir) and ref_dir not in m.__path__:
                m.__path__.append(ref_dir)
    for mod in _MODULES:
        m = importlib.import_module(f"playslot_b200.models.{mod}")
        sys.modules[f"models.{mod}"] = m
        names.append(f"models.{mod}")
    if reference_src:
        ref_models = os.path.join(reference_src, "models")
        if os.path.isdir(ref_models) and ref_models not in mirror.__path__:
            mirror.__path__.append(ref_models)
    return names


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m playslot_b200.launch <reference script.py> [script args]")
    script = os.path.abspath(argv[0])
    src = os.path.dirname(script)
    sys.path.insert(0, src)
    install(reference_src=src)
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
