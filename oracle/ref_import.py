"""Import the UNMODIFIED reference modules from /root/reference (this container only).
TEST INFRASTRUCTURE: used by oracle/validate_against_reference.py and
tests/golden/make_golden.py.  /root/reference does not exist on the GPU box; nothing in the
gpu tests / smoke / bench imports this file."""
import importlib
import importlib.util
import os
import pickle
import sys
import types

REF = os.environ.get("EMO_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    return os.path.isdir(os.path.join(REF, "stage2_accompaniment"))


def _load_pkg(name, path):
    """Load directory ``path`` as top-level package ``name`` (fresh copy)."""
    for k in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
        del sys.modules[k]
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(path, "__init__.py"), submodule_search_locations=[path])
    if spec is None or not os.path.exists(os.path.join(path, "__init__.py")):
        mod = types.ModuleType(name)
        mod.__path__ = [path]
        sys.modules[name] = mod
        return mod
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def stage1_model():
    _load_pkg("ref_s1_model", os.path.join(REF, "stage1_compose", "model"))
    return importlib.import_module("ref_s1_model.plain_transformer")


def stage2_gpt2():
    """music_gpt2.py under HF transformers 5.5.0 with a two-part shim that restores the
    README-pinned 4.28.0 behaviour of ``GPT2Block.forward`` (music_gpt2.py:86 relies on both):
      (a) 4.28.0 returned a tuple -> re-wrap the Tensor return;
      (b) 4.28.0's GPT2Attention applied its own causal ``bias`` buffer (tril, masked with
          finfo.min); 5.x standalone blocks attend bidirectionally unless an attention_mask is
          passed [probe: prefix-invariance fails unshimmed] -> inject the additive causal mask."""
    import torch
    from transformers.models.gpt2 import modeling_gpt2
    if not getattr(modeling_gpt2.GPT2Block.forward, "_emo_shim", False):
        orig = modeling_gpt2.GPT2Block.forward

        def fwd(self, hidden_states, *a, **k):
            if not a and k.get("attention_mask") is None:
                L = hidden_states.shape[1]
                neg = torch.finfo(hidden_states.dtype).min
                mask = torch.full((L, L), neg, dtype=hidden_states.dtype, device=hidden_states.device).triu(1)
                k["attention_mask"] = mask[None, None]
            out = orig(self, hidden_states, *a, **k)
            return out if isinstance(out, tuple) else (out,)
        fwd._emo_shim = True
        modeling_gpt2.GPT2Block.forward = fwd
    _install_standin()
    _load_pkg("ref_s2_model", os.path.join(REF, "stage2_accompaniment", "model"))
    return importlib.import_module("ref_s2_model.music_gpt2")


def _install_standin():
    root = os.path.dirname(_HERE)
    if root not in sys.path:
        sys.path.insert(0, root)
    standin = os.path.join(_HERE, "standin")
    if standin not in sys.path:
        sys.path.insert(0, standin)


def stage2_performer():
    _install_standin()
    _load_pkg("ref_s2_model", os.path.join(REF, "stage2_accompaniment", "model"))
    return importlib.import_module("ref_s2_model.music_performer")


def sampling_functions():
    """temperature / nucleus of both stages without importing the scripts' heavy deps."""
    out = {}
    for stage, rel in (("stage2", "stage2_accompaniment/inference.py"),
                       ("stage1", "stage1_compose/inference_utils.py")):
        src = open(os.path.join(REF, rel)).read()
        import ast
        tree = ast.parse(src)
        keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("temperature", "nucleus")]
        mod = ast.Module(body=keep, type_ignores=[])
        import numpy as np
        import scipy, scipy.special
        ns = {"np": np, "scipy": scipy}
        exec(compile(mod, rel, "exec"), ns)
        out[stage] = (ns["temperature"], ns["nucleus"])
    return out


def decode_functions():
    """The reference decode loops, extracted verbatim (AST) from the unmodified scripts without
    importing their heavy / missing dependencies (dataloader, convert2midi, miditoolkit, pickle5):
      stage2: generate_conditional, temperature, nucleus, get_position_idx   (inference.py:71-100,206-327)
      stage1: generate_plain_xl, temperature, nucleus, match_emotion_key       (inference_utils.py:14-143)
    Returns {'stage1': namespace, 'stage2': namespace}; namespaces are plain dicts whose `nucleus` can be
    swapped (e.g. for argmax) before calling the loop."""
    import ast
    import time
    import numpy as np
    import scipy, scipy.special
    import torch
    out = {}
    wanted = {"stage2": ("stage2_accompaniment/inference.py",
                         ("generate_conditional", "temperature", "nucleus", "get_position_idx")),
              "stage1": ("stage1_compose/inference_utils.py",
                         ("generate_plain_xl", "temperature", "nucleus", "get_position_idx", "match_emotion_key"))}
    for stage, (rel, names) in wanted.items():
        tree = ast.parse(open(os.path.join(REF, rel)).read())
        keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
        ns = {"np": np, "scipy": scipy, "torch": torch, "time": time, "max_dec_inp_len": 2048,
              "tensor_to_numpy": lambda t: t.cpu().detach().numpy(),
              "MAJOR_KEY": np.array(['C', 'C#', 'D', 'D#', 'E', 'F', 'F#', 'G', 'G#', 'A', 'A#', 'B']),
              "MINOR_KEY": np.array(['c', 'c#', 'd', 'd#', 'e', 'f', 'f#', 'g', 'g#', 'a', 'a#', 'b'])}
        exec(compile(ast.Module(body=keep, type_ignores=[]), rel, "exec"), ns)
        out[stage] = ns
    return out
