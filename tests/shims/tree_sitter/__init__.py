"""tree_sitter stand-in: the names the reference chunker imports.  No grammar is available, so
`TreeSitterChunker.chunk_file` takes its own documented fallback (plain line / character segmentation,
tree_sitter_chunker.py:95-104)."""


class Language:
    pass


class Node:
    pass


class Parser:
    def set_language(self, language):
        raise RuntimeError("no tree-sitter grammar in the test image")

    def parse(self, source):
        raise RuntimeError("no tree-sitter grammar in the test image")
