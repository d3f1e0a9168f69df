from typing import NewType

ReplayItemID = NewType("ReplayItemID", int)  # slimdqn/sample_collection/__init__.py:1-3
