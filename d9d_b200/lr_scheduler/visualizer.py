from __future__ import annotations

from collections.abc import Callable

from torch import nn
from torch.optim import SGD, Optimizer
from torch.optim.lr_scheduler import LRScheduler

SchedulerFactory = Callable[[Optimizer], LRScheduler]


def lr_history(factory: SchedulerFactory, num_steps: int, init_lr: float = 1.0) -> list[float]:
    """Simulate ``num_steps`` scheduler steps on a dummy optimizer and return the learning rate of each step."""
    optimizer = SGD(nn.Linear(1, 1).parameters(), lr=init_lr)
    scheduler = factory(optimizer)
    history = []
    for _ in range(num_steps):
        history.append(optimizer.param_groups[0]["lr"])
        scheduler.step()
    return history


def visualize_lr_scheduler(factory: SchedulerFactory, num_steps: int, init_lr: float = 1.0) -> None:
    """Interactive plot of the schedule (needs the optional ``plotly`` dependency)."""
    try:
        import plotly.graph_objects as go
    except ImportError as exc:
        raise ImportError("You have to install `plotly` dependency to use scheduler visualization") from exc
    lrs = lr_history(factory, num_steps, init_lr)
    fig = go.Figure(go.Scatter(x=list(range(num_steps)), y=lrs, mode="lines", name="Learning Rate"))
    fig.update_layout(title="Scheduler", xaxis_title="Steps", yaxis_title="Learning Rate", template="plotly_white",
                      hovermode="x unified", height=500)
    fig.show()
