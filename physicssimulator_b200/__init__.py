"""physicssimulator_b200 — B200-native batched multi-world stepper behind the API of
Mic43/PhysicsSimulator's PhysicsSimulatorCore."""
