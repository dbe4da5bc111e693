"""Per-step agent outputs recorded in the sample buffer (rlpyt/agents/pg/base.py:4)."""
from rlpyt_b200.utils.collections import namedarraytuple

AgentInfo = namedarraytuple("AgentInfo", ["dist_info", "value"])
AgentInfoRnn = namedarraytuple("AgentInfoRnn", ["dist_info", "value", "prev_rnn_state"])   # pg/base.py:5-6
