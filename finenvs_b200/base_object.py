"""Mirror of the reference's finenvs/base_object.py:7-9 (every reference class exposes .print())."""
from pprint import pprint


class BaseObject(object):
    def print(self) -> None:
        pprint(vars(self))
