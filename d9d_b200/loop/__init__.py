"""Training / inference orchestration: configs, provider protocols, event bus, components, runners."""
