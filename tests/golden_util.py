"""Helpers turning the committed golden fixtures into WindowBatches."""
from hypo_b200.batch import WindowSpec, build_batch


def group_batch(group):
    specs = [WindowSpec(w["draft"], w["internal"], w["pre"], w["suf"], w["n_empty"], w["wtype"])
             for w in group["windows"]]
    return build_batch(specs), [w["consensus"] for w in group["windows"]]
