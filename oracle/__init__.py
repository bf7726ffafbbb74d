"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference forward and the HuggingFace oracle.

Nothing under vit.triton_b200/ imports this package.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may use it, and only as the checker or the timed
CPU baseline — never as the product path.
"""
