"""Generate tests/golden/keras_names.json: for every fixture case, the Keras weight name (`<layer>/<weight>:0`, what
`layer.weights[i].name` / the HDF5 `weight_names` attribute hold) of every canonical weight, obtained by EXECUTING the
reference's model.py / resnet.py on `minikeras` with Keras' layer auto-naming rule (snake_case class name + per-prefix
uid from 1; Bidirectional renames its copies forward_<name> / backward_<name>) -- i.e. the creation order is the
reference's, the naming rule is Keras' documented one.  TEST INFRASTRUCTURE ONLY; build container only.

    python tests/golden/make_golden_names.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg                                  # noqa: E402  (imports the reference on minikeras)
import minikeras as mk                                    # noqa: E402


def names_of_case(case):
    mg.run_case(case)
    out = {}
    for cname, layer in mk.USED:
        parts = cname.split("/")
        weight = parts[-1]
        if len(parts) >= 3 and parts[-2] in ("forward", "backward"):         # Bidirectional(CuDNNGRU)
            out[cname] = "%s/%s/%s:0" % ("/".join(parts[:-2]), layer.keras_name, weight)
        else:
            out[cname] = "%s/%s:0" % (layer.keras_name, weight)
    return out


def main():
    import contextlib, io
    res = {}
    for case in mg.CASES:
        with contextlib.redirect_stdout(io.StringIO()):
            res[case] = {"T": mg.CASES[case][0], "kwargs": mg.CASES[case][4], "names": names_of_case(case)}
    path = os.path.join(HERE, "keras_names.json")
    with open(path, "w") as f:
        json.dump(res, f, indent=0, sort_keys=True)
    print("wrote", path, {k: len(v["names"]) for k, v in res.items()})


if __name__ == "__main__":
    main()
