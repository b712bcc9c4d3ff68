"""CPU oracle for the qMC diagram-evaluation hot path — TEST INFRASTRUCTURE ONLY.

See qiw_oracle.cpp for the restatement and its parity status (pinned against the reference's
golden vectors).  Nothing in the product package imports this.
"""
