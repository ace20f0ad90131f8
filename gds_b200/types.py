"""Enums and base interfaces (gds/types.py), plus the fixed-step members `Integrators.rk4` / `Integrators.euler`."""
from __future__ import annotations

from abc import ABC, abstractmethod
from enum import Enum
from typing import Callable

import numpy as np


class GraphDomain(Enum):
    nodes = 0
    edges = 1
    triangles = 2
    faces = 3


class IterationMode(Enum):
    none = 0
    dydt = 1
    cvx = 2
    map = 3
    traj = 4
    nil = 5


class Integrators(Enum):
    rk45 = 0
    lsoda = 1
    dop853 = 2
    implicit_midpoint = 3
    rk4 = 4      # fixed-step classical Runge-Kutta, device resident
    euler = 5    # fixed-step forward Euler, device resident


class Steppable(ABC):
    def __init__(self, iter_mode: IterationMode):
        self.iter_mode = iter_mode

    @abstractmethod
    def step(self, dt: float):
        ...

    @abstractmethod
    def reset(self):
        ...


class Observable(ABC):
    """something with a domain X (point -> index), a time t and values y (gds/types.py:66-105)"""

    def __init__(self, X):
        self.X = X
        self._iX = None
        self.ndim = len(X)

    @property
    def iX(self):
        if self._iX is None:
            self._iX = {i: x for x, i in self.X.items()}
        return self._iX

    @property
    @abstractmethod
    def t(self) -> float:
        ...

    @property
    @abstractmethod
    def y(self) -> np.ndarray:
        ...

    def __getitem__(self, idx):
        return self.y.__getitem__(idx)

    def __call__(self, x):
        return self.y[self.X[x]]

    def __len__(self):
        return self.ndim

    def project(self, other: type, view: Callable, *args, **kwargs):
        """derived read-only observable computed from this one (types.py:95-105)"""
        parent = self

        class Projection(other):
            def __init__(self_):
                other.__init__(self_, *args, **kwargs)

            @property
            def t(self_):
                return parent.t

            @property
            def y(self_):
                from .fields import evaluate_view      # device operators + reductions where the view allows it
                return evaluate_view(view, parent)
        return Projection()


class PointObservable(Observable):
    def __init__(self, retention=600, min_res=1e-3, **kwargs):
        self.render_params = {'retention': retention, 'min_res': min_res, **kwargs}
        super().__init__(dict())


class VectorObservable(Observable):
    def __init__(self, domain, retention=240, **kwargs):
        self.domain = [str(x) for x in domain]
        self.render_params = {'retention': retention, **kwargs}
        super().__init__(dict())
        self.ndim = len(domain)
