"""Shared helpers for the tests (oracle access lives here: tests are allowed to use it)."""
import numpy as np
import torch

from tests.golden import cases as G


def rms(a):
    a = np.asarray(a, np.float64)
    return float(np.sqrt(np.mean(np.square(a)))) if a.size else 0.0


def load_golden(name):
    z = np.load(G.path(name), allow_pickle=False)
    return {k: z[k] for k in z.files}


def check_inputs_match_golden(name, inp, gold):
    """The goldens were produced from seeded inputs; make sure this torch build regenerates
    the same inputs (otherwise the comparison is meaningless)."""
    for k, v in G.input_checksums(inp).items():
        assert k in gold, k
        assert abs(float(gold[k]) - v) <= 1e-9 * max(1.0, abs(v)), \
            "input %s of %s differs from the one the golden was made with" % (k, name)


def port_outputs(name, inp):
    """Run the oracle's torch port on a golden case."""
    from oracle import torch_port as tp
    case = inp["case"]
    with torch.no_grad():
        if case["kind"] == "sins":
            return tp.sins_forward(inp["f0"], inp["ctrls"], G.SR, G.P, noise=inp["noise"],
                                   initial_phase=inp.get("initial_phase"))
        if case["kind"] == "combsub":
            return tp.combsub_forward(inp["f0"], inp["ctrls"], G.SR, G.P, noise=inp["noise"])
        if case["kind"] == "superfast":
            return tp.superfast_forward(inp["f0"], inp["ctrls"], G.SR, G.P, case["win"], noise=inp["noise"])
        if case["kind"] == "combsubfast":
            return tp.combsubfast_forward(inp["f0"], inp["ctrls"], G.SR, G.P, noise=inp["noise"],
                                          initial_phase=inp.get("initial_phase"))
        if case["kind"] == "source_module":
            gold = load_golden(name)          # the seeded Linear(9 -> 1) parameters travel with the golden
            return tp.source_module_forward(inp["f0"], case["upp"], G.SR, torch.from_numpy(gold["weight"]),
                                            torch.from_numpy(gold["bias"]), case["harmonic_num"],
                                            rand_ini=inp["rand_ini"], noise=inp["noise"])
        return tp.sinegen_forward(inp["f0"], case["upp"], G.SR, case["harmonic_num"],
                                  rand_ini=inp["rand_ini"], noise=inp["noise"])


def closed_form_outputs(name, inp):
    from oracle import closed_form as cf
    case = inp["case"]
    npc = lambda d: {k: v.numpy() for k, v in d.items()}
    if case["kind"] == "sins":
        ip = inp.get("initial_phase")
        return cf.sins(inp["f0"].numpy(), npc(inp["ctrls"]), G.SR, G.P, inp["noise"].numpy(),
                       None if ip is None else ip.numpy())
    if case["kind"] == "combsub":
        return cf.combsub(inp["f0"].numpy(), npc(inp["ctrls"]), G.SR, G.P, inp["noise"].numpy())
    if case["kind"] == "combsubfast":
        ip = inp.get("initial_phase")
        return cf.combsubfast(inp["f0"].numpy(), npc(inp["ctrls"]), G.SR, G.P, inp["noise"].numpy(),
                              None if ip is None else ip.numpy())
    if case["kind"] == "superfast":
        return cf.superfast(inp["f0"].numpy(), npc(inp["ctrls"]), G.SR, G.P, case["win"], inp["noise"].numpy())
    sines = cf.sinegen(inp["f0"].numpy(), case["upp"], G.SR, inp["rand_ini"].numpy().reshape(-1), inp["noise"].numpy())
    if case["kind"] == "source_module":
        gold = load_golden(name)
        return {"out": np.tanh(sines @ gold["weight"].astype(np.float64).T + gold["bias"].astype(np.float64))}
    return {"out": sines}
