"""Event <-> vocabulary-index codec: the contract of the reference's event_codec.py:34-112.

Shift events occupy the first block [0, max_shift_steps]; every other event type gets the
next contiguous block in declaration order.  Implemented with a precomputed offset table.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Tuple


@dataclasses.dataclass
class EventRange:
    type: str
    min_value: int
    max_value: int


@dataclasses.dataclass
class Event:
    type: str
    value: int


class Codec:
    """Encode and decode events (event_codec.py:34-112)."""

    def __init__(self, max_shift_steps: int, steps_per_second: float, event_ranges: List[EventRange]):
        self.steps_per_second = steps_per_second
        self._shift_range = EventRange(type='shift', min_value=0, max_value=max_shift_steps)
        self._event_ranges = [self._shift_range] + list(event_ranges)
        names = [er.type for er in self._event_ranges]
        assert len(names) == len(set(names)), 'event types must be unique'
        # type -> (first index, EventRange); plus a sorted list of block starts for decoding
        self._table: Dict[str, Tuple[int, EventRange]] = {}
        start = 0
        for er in self._event_ranges:
            self._table[er.type] = (start, er)
            start += er.max_value - er.min_value + 1
        self._num_classes = start

    @property
    def num_classes(self) -> int:
        return self._num_classes

    def is_shift_event_index(self, index: int) -> bool:
        return (self._shift_range.min_value <= index) and (index <= self._shift_range.max_value)

    @property
    def max_shift_steps(self) -> int:
        return self._shift_range.max_value

    def encode_event(self, event: Event) -> int:
        """Encode an event to an index."""
        if event.type not in self._table:
            raise ValueError(f'Unknown event type: {event.type}')
        start, er = self._table[event.type]
        if not er.min_value <= event.value <= er.max_value:
            raise ValueError(f'Event value {event.value} is not within valid range '
                             f'[{er.min_value}, {er.max_value}] for type {event.type}')
        return start + event.value - er.min_value

    def event_type_range(self, event_type: str) -> Tuple[int, int]:
        """Return [min_id, max_id] for an event type."""
        if event_type not in self._table:
            raise ValueError(f'Unknown event type: {event_type}')
        start, er = self._table[event_type]
        return start, start + (er.max_value - er.min_value)

    def decode_event_index(self, index: int) -> Event:
        """Decode an event index to an Event."""
        for start, er in self._table.values():
            if start <= index <= start + er.max_value - er.min_value:
                return Event(type=er.type, value=er.min_value + index - start)
        raise ValueError(f'Unknown event index: {index}')
