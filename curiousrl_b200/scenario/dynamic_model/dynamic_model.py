"""Property bag every model-based scenario exposes to the iLQR solver.

Same read-only surface as the reference's `DynamicModelWrapper`
(CuriousRL/scenario/dynamic_model/dynamic_model.py:55-101): `dynamic_function`
(sympy Array, x_{t+1} = f(x_t, u_t)), `x_u_var`, `constr` [n+m, 2],
`init_state` [n,1], `init_action` [T,m,1], `obj_fun` (sympy scalar),
`add_param_var`, `add_param` [T,p], `n`, `m`, `T`, `name`.  `name` is the class
name and is the key the B200 solver uses to pick a precompiled device model.

Plotting (`create_plot`, dynamic_model.py:36) needs matplotlib, which is a GUI
concern outside the hot path; it is imported lazily and only if asked for.
"""
from __future__ import annotations

from ..scen_wrapper import ScenarioWrapper


class DynamicModelWrapper(ScenarioWrapper):
    def __init__(self, dynamic_function, x_u_var, constr, init_state, init_action,
                 obj_fun, add_param_var=None, add_param=None):
        self._dynamic_function = dynamic_function
        self._x_u_var = x_u_var
        self._constr = constr
        self._init_state = init_state
        self._init_action = init_action
        self._obj_fun = obj_fun
        self._add_param_var = add_param_var
        self._add_param = add_param
        self._n = int(init_state.shape[0])
        self._m = int(len(x_u_var)) - self._n
        self._T = int(init_action.shape[0])
        self._fig = None
        self._ax = None

    # capability flags (dynamic_model.py:46-53)
    def with_model(self):
        return True

    def is_action_discrete(self):
        return False

    def is_output_image(self):
        return False

    dynamic_function = property(lambda self: self._dynamic_function)
    x_u_var = property(lambda self: self._x_u_var)
    constr = property(lambda self: self._constr)
    init_state = property(lambda self: self._init_state)
    init_action = property(lambda self: self._init_action)
    obj_fun = property(lambda self: self._obj_fun)
    add_param_var = property(lambda self: self._add_param_var)
    add_param = property(lambda self: self._add_param)
    n = property(lambda self: self._n)
    m = property(lambda self: self._m)
    T = property(lambda self: self._T)

    @property
    def name(self):
        return type(self).__name__

    def create_plot(self, figsize=(5, 5), xlim=(-6, 6), ylim=(-6, 6)):
        import matplotlib.pyplot as plt  # optional dependency, GUI only
        if self._fig is None:
            self._fig = plt.figure(figsize=figsize)
            self._ax = self._fig.add_subplot(111)
            self._ax.axis("equal")
            self._ax.set_xlim(*xlim)
            self._ax.set_ylim(*ylim)
            self._ax.grid(True)
        return self._fig, self._ax

    def play(self, logger_folder=None, no_iter=-1):
        raise NotImplementedError
