"""Application registrations (mirror of /root/reference config/app_registration.py:1-5, which
instantiates the registry and registers nothing).  Register B200 retrieval applications here::

    from rag_arc_b200.configs import HybridRetrieverConfig
    registrator.register("configs/hybrid.json", "hybrid_search", HybridRetrieverConfig)
    retriever = registrator.get_object("hybrid_search")
"""
from ..framework.register import Register

registrator = Register()
