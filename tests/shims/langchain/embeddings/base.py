class Embeddings:
    """Interface the reference type-annotates against (langchain.embeddings.base.Embeddings)."""

    def embed_documents(self, texts):
        raise NotImplementedError

    def embed_query(self, text):
        raise NotImplementedError
