"""CPU oracle for the IVF_FLAT hot path -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference (rmontanana/semcode) holds no golden vector, known-answer
test or fixture for this path and its engine (pymilvus 2.6.2 ->
Milvus v2.4.4 -> knowhere -> FAISS IndexIVFFlat) is neither under /root/reference nor
installable offline.  The oracle is therefore a restatement of the *published* FAISS
IndexIVFFlat algorithm, anchored on the reference's own call sites
(src/semcode/storage/milvus_store.py:76-83 index spec, :141-147 search params).

Nothing under semcode_b200/ may import this package.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs use it, and only as the checker or the
timed CPU baseline -- never as the product path.
"""
