"""Ground-truth simulator on the device: E[Y | do(X = x)] of a graph's structural equation model by Monte Carlo
(cbo_sem_eval, csrc/sem.cu) -- the device form of compute_interventions (reference graph_functions.py:48-77).

A graph describes its SEM as a program (`graph.device_sem()`):
    {"noise": [names of the per-sample random inputs],
     "nodes": [(name, constant, [(source name, func, coef, scale), ...]), ...]      # topological order
     "draw":  callable(num_samples, intervened_names, seed) -> (len(noise), num_samples) float64 array}
`draw` replays the HOST function's use of the seeded NumPy stream (np.random.seed(seed); randn(num_samples, nodes); then
whatever the un-intervened stochastic nodes draw), so device and host see bit-identical noise.  compute_interventions
reseeds with seed 1 on every call: the noise of one intervention signature is drawn once and kept in HBM."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import _lib


class DeviceSEM:
    def __init__(self, graph, num_samples: int = 100000, device="cuda:0", target: str = "Y", seed: int = 1):
        if not torch.cuda.is_available():
            raise RuntimeError("DeviceSEM needs a CUDA device (use graph_functions.compute_interventions on the host)")
        self.lib = _lib.load()
        self.dev = torch.device(device)
        self.num_samples, self.seed = int(num_samples), int(seed)
        prog = graph.device_sem()
        self.noise_names: List[str] = list(prog["noise"])
        self.node_names: List[str] = [n[0] for n in prog["nodes"]]
        self.draw = prog["draw"]
        index = {name: i for i, name in enumerate(self.noise_names)}
        index.update({name: len(self.noise_names) + i for i, name in enumerate(self.node_names)})
        nodes = (_lib.SemNode * len(prog["nodes"]))()
        terms: List[Tuple[int, int, float, float]] = []
        for i, (name, constant, tt) in enumerate(prog["nodes"]):
            nodes[i].first_term, nodes[i].num_terms, nodes[i].constant = len(terms), len(tt), float(constant)
            for src, func, coef, scale in tt:
                if index[src] >= len(self.noise_names) + i:
                    raise ValueError(f"SEM program is not in topological order at node {name!r} (source {src!r})")
                terms.append((index[src], _lib.SEM_FUNCS[func], float(coef), float(scale)))
        if len(nodes) > _lib.CBO_SEM_MAX_NODES or len(terms) > _lib.CBO_SEM_MAX_TERMS:
            raise ValueError("SEM program too large for cbo_sem_eval")
        tarr = (_lib.SemTerm * max(len(terms), 1))()
        for k, (a, f, c, s) in enumerate(terms):
            tarr[k].src, tarr[k].func, tarr[k].coef, tarr[k].scale = a, f, c, s
        self.num_nodes, self.num_terms = len(nodes), len(terms)
        self.d_nodes = torch.frombuffer(bytearray(bytes(nodes)), dtype=torch.uint8).to(self.dev)
        self.d_terms = torch.frombuffer(bytearray(bytes(tarr)), dtype=torch.uint8).to(self.dev)
        self.target = self.node_names.index(target)
        self._noise: Dict[tuple, torch.Tensor] = {}

    def _noise_for(self, intervened: Sequence[str]) -> torch.Tensor:
        key = tuple(sorted(intervened))
        if key not in self._noise:
            state = np.random.get_state()       # the host function leaves the global stream reseeded; do not disturb callers here
            try:
                arr = np.ascontiguousarray(self.draw(self.num_samples, list(intervened), self.seed), dtype=np.float64)
            finally:
                np.random.set_state(state)
            assert arr.shape == (len(self.noise_names), self.num_samples)
            self._noise[key] = torch.from_numpy(arr).to(self.dev)
        return self._noise[key]

    def mean_target(self, variables: Sequence[str], X: np.ndarray) -> np.ndarray:
        """Monte-Carlo mean of the target under do(variables = X[b]) for every row b of X (B, d) -> (B,)."""
        X = np.ascontiguousarray(np.asarray(X, np.float64).reshape(-1, len(variables)))
        B = X.shape[0]
        mask = np.zeros((B, self.num_nodes), np.int32)
        val = np.zeros((B, self.num_nodes), np.float64)
        for k, v in enumerate(variables):
            mask[:, self.node_names.index(v)] = 1
            val[:, self.node_names.index(v)] = X[:, k]
        noise = self._noise_for(variables)
        d_mask, d_val = torch.from_numpy(mask).to(self.dev), torch.from_numpy(val).to(self.dev)
        partials = torch.empty((B * _lib.CBO_SEM_BLOCKS,), dtype=torch.float64, device=self.dev)
        mean = torch.empty((B,), dtype=torch.float64, device=self.dev)
        st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        _lib.check(self.lib.cbo_sem_eval(C.c_void_p(self.d_nodes.data_ptr()), self.num_nodes, C.c_void_p(self.d_terms.data_ptr()),
                                         self.num_terms, C.c_void_p(noise.data_ptr()), len(self.noise_names), self.num_samples,
                                         C.c_void_p(d_mask.data_ptr()), C.c_void_p(d_val.data_ptr()), B, self.target,
                                         C.c_void_p(partials.data_ptr()), C.c_void_p(mean.data_ptr()), st), "cbo_sem_eval")
        return mean.cpu().numpy()
