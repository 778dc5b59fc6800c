"""Every reference class exposes `.print()` (finenvs/base_object.py:7-9): a dump of the instance's attributes."""
import pprint as _pprint


class BaseObject:
    def print(self) -> None:
        _pprint.pprint(self.__dict__)
