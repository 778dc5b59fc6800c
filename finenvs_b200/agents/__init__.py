"""Callers on either side of the env step: rollout storage for the reference's agents.
The agents themselves (finenvs/agents/*) are consumers and stay the reference's own code."""
