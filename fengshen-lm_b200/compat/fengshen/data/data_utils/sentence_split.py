import re

_END = re.compile(r"(？|。|[！]+|!|\…\…)")   # the reference's sentence terminators (sentence_split.py:24)


class ChineseSentenceSplitter(object):
    """fengshen/data/data_utils/sentence_split.py:4-35: cut a paragraph after each terminator, move a closing quote that starts
    a sentence back to the previous one, then glue the pieces of a quotation that was cut in the middle."""

    def merge_symmetry(self, sentences, symmetry=('“', '”')):
        opener, closer = symmetry
        out, inside = [], False
        for s in sentences:
            has_open, has_close = opener in s, closer in s
            if has_open and not has_close:
                inside = True
                out.append(s)
            elif inside:   # still inside the quotation, or its closing mark arrives: glue onto the running piece
                out[-1] += s
                if has_close:
                    inside = False
            else:
                out.append(s)
        return [s.strip() for s in out if s.strip()]

    def to_sentences(self, paragraph):
        parts = _END.split(paragraph)           # text, terminator, text, terminator, ..., text
        parts.append("")
        pieces = [(parts[i] + parts[i + 1]).strip() for i in range(0, len(parts) - 1, 2)]
        pieces = [p for p in pieces if p]
        for j in range(1, len(pieces)):
            if pieces[j][0] == '”':
                pieces[j - 1] += '”'
                pieces[j] = pieces[j][1:]
        return self.merge_symmetry(pieces)

    def tokenize(self, text):
        return self.to_sentences(text)
