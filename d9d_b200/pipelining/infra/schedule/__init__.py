"""Schedule namespace in the reference's layout (``d9d/pipelining/infra/schedule``).

The implementation lives one level up (``infra/{action,simulator,programs,communications,topology,executor}.py``): every
schedule here is a *policy* run through one event-driven simulator instead of a hand-written program per schedule.
These sub-packages re-export it under the import paths the reference's users know.
"""
