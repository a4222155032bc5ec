import os

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSET_DIR = os.path.join(REPO_ROOT, "asset")
INCLUDE_DIR = os.path.join(REPO_ROOT, "include")
COMPILER = os.path.join(REPO_ROOT, "OpenABL")
RUNTIME_LIB = os.path.join(ASSET_DIR, "cuda", "libabl_cuda.so")
MODEL_CACHE = os.path.join(REPO_ROOT, "build", "models")
EXAMPLES_DIR = os.path.join(REPO_ROOT, "examples")
