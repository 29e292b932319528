"""
Recipe that makes the UNMODIFIED reference importable on the GPU box.  TEST / BASELINE INFRASTRUCTURE.

The reference is pure Python: "building" it means copying the few source files of the slot pipeline from where they lie
under /root/reference into oracle/_ref/ (git-ignored, NOT gpurun-ignored, so it travels with the snapshot like our own
built .so; nothing of it enters the history).  Only bench.py's baseline legs and tests/ import it, as the checker / the
timed baseline -- never the product path.

    python oracle/install_ref.py          # copies; prints the file count
    from oracle.install_ref import load_reference
    setup_model, setup_predictor, experiment_json = load_reference()

What is copied: src/models (the nn.Modules of the path), src/lib/{setup_model, logger, utils, schedulers,
config, loss}.py, src/configs, src/CONFIG.py and the shipped experiments/*/predictors/*/experiment_params.json.
"""
import glob
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
REF = os.environ.get("PLAYSLOT_REFERENCE", "/root/reference")
LIB_FILES = ("setup_model.py", "logger.py", "utils.py", "schedulers.py", "config.py", "loss.py")


def available():
    return os.path.isfile(os.path.join(DEST, "src", "lib", "setup_model.py"))


def install(force=False):
    """ copy the reference sources of the path into oracle/_ref; returns the number of files (0 when the reference tree is absent) """
    if not os.path.isdir(os.path.join(REF, "src", "models")):
        return 0
    if available() and not force:
        return sum(len(f) for _, _, f in os.walk(DEST))
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    ign = shutil.ignore_patterns("__pycache__", "*.pyc")
    shutil.copytree(os.path.join(REF, "src", "models"), os.path.join(DEST, "src", "models"), ignore=ign)
    shutil.copytree(os.path.join(REF, "src", "configs"), os.path.join(DEST, "src", "configs"), ignore=ign)
    os.makedirs(os.path.join(DEST, "src", "lib"), exist_ok=True)
    for f in LIB_FILES:
        shutil.copy2(os.path.join(REF, "src", "lib", f), os.path.join(DEST, "src", "lib", f))
    shutil.copy2(os.path.join(REF, "src", "CONFIG.py"), os.path.join(DEST, "src", "CONFIG.py"))
    for j in glob.glob(os.path.join(REF, "experiments", "*", "predictors", "*", "experiment_params.json")):
        rel = os.path.relpath(j, REF)
        os.makedirs(os.path.dirname(os.path.join(DEST, rel)), exist_ok=True)
        shutil.copy2(j, os.path.join(DEST, rel))
    with open(os.path.join(DEST, "README"), "w") as f:
        f.write("Unmodified copy of the chrono-motz/PlaySlot sources of the slot pipeline, made by oracle/install_ref.py.\n"
                "Baseline / test infrastructure only; git-ignored.\n")
    return sum(len(f) for _, _, f in os.walk(DEST))


_loaded = None


def load_reference():
    """ (setup_model, setup_predictor, experiment_json(name, predictor='PlaySlot')) of the unmodified reference in oracle/_ref.
    CONFIG.py:11 resolves its paths from the working directory at import time, so the import happens with cwd = oracle/_ref. """
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise FileNotFoundError(f"{DEST} is missing: run `python oracle/install_ref.py` where /root/reference is mounted")
    cwd = os.getcwd()
    sys.dont_write_bytecode = True
    os.chdir(DEST)
    try:
        sys.path.insert(0, os.path.join(DEST, "src"))
        from lib.setup_model import setup_model, setup_predictor
    finally:
        os.chdir(cwd)

    def experiment_json(name, predictor="PlaySlot"):
        return json.load(open(os.path.join(DEST, "experiments", name, "predictors", predictor, "experiment_params.json")))

    # get_available_configs (setup_model.py:45,79) lists CONFIG["paths"]["configs_path"], an absolute path fixed at import
    _loaded = (setup_model, setup_predictor, experiment_json)
    return _loaded


if __name__ == "__main__":
    n = install(force="--force" in sys.argv)
    print(f"oracle/_ref: {n} files" if n else f"reference tree {REF} not found; nothing installed")
