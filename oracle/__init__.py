"""CPU oracle of the hybrid-ensemble inference path (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (oct_segmentation_b200/, src/) never does.
"""
