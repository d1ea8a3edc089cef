"""Name-keyed lookup in PARESIS's parameter files (xmlFiles/*.xml), via xml.dom.minidom.

The reference scans ``<experiment>/<sample>/<source>/<detector>`` blocks for a ``<name>``
match and reads child nodes by tag (Experiment.py:159-197, Sample.py:44-77, Source.py:49-77,
Detector.py:55-76); optional settings are detected by their presence among the children.
"""
from xml.dom import minidom


def text_of(node):
    return node.childNodes[0].nodeValue


class Entry:
    """One named block; ``get(tag)`` reads the first descendant with that tag."""

    def __init__(self, element):
        self.element = element

    def has(self, tag):
        return any(child.localName == tag for child in self.element.childNodes)

    def get(self, tag, cast=str):
        return cast(text_of(self.element.getElementsByTagName(tag)[0]))


def find_entry(xml_path, block_tag, name):
    """Return the block whose <name> equals ``name`` or None."""
    doc = minidom.parse(xml_path)
    for element in doc.documentElement.getElementsByTagName(block_tag):
        if text_of(element.getElementsByTagName("name")[0]) == name:
            return Entry(element)
    return None
